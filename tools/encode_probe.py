"""Encode (predict) throughput of the level chain for a few chunk sizes.  RQK_ENC_CHUNK is read once per process, so
each size runs in its own process: python tools/encode_probe.py [rows]"""
import os, subprocess, sys
if len(sys.argv) > 1 and sys.argv[1] == "child":
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import torch
    from generative_ranking_recommender_b200 import engine
    n = int(sys.argv[2]); dev = torch.device("cuda:0")
    g = torch.Generator(device=dev); g.manual_seed(1)
    x = torch.randn((n, 512), device=dev, generator=g)
    cs = [x[torch.randperm(n, device=dev)[:k]].contiguous() * s for k, s in ((128, 1.0), (128, 0.05), (256, 0.05))]
    for mode in (0, 1):
        for _ in range(2):
            ids = engine.encode(x, cs, [128, 128, 256], [512], mode=mode)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            ids = engine.encode(x, cs, [128, 128, 256], [512], mode=mode)
        e1.record(); torch.cuda.synchronize()
        t = e0.elapsed_time(e1) / 3
        print(f"chunk {os.environ.get('RQK_ENC_CHUNK', 'default'):>8s} mode {mode}: {t:7.2f} ms for {n} rows = {n / t / 1e3:7.1f} M vectors/s "
              f"(X stream alone: {n * 2048 / 6551.4e9 * 1e3:.2f} ms)", flush=True)
else:
    n = sys.argv[1] if len(sys.argv) > 1 else "1000000"
    for c in ("18944", "37888", "75776", "151552", "1048576"):
        subprocess.run([sys.executable, __file__, "child", n], env=dict(os.environ, RQK_ENC_CHUNK=c))
