"""Timing attribution of the score pass (RQK_SCORE_DEBUG variants produce WRONG results on purpose)."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
code = r'''
import sys, torch
sys.path.insert(0, %r)
from generative_ranking_recommender_b200 import engine
dev = torch.device("cuda:0")
for n, k in ((1000000, 128), (1000000, 256)):
    x = torch.randn((n, 512), device=dev); c = x[:k].clone()
    for scores in (True, False):
        for _ in range(2): engine.score_pass(x, c, scores=scores, argmin=True, counts=True)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(5): engine.score_pass(x, c, scores=scores, argmin=True, counts=True)
        b.record(); torch.cuda.synchronize()
        print("k", k, "scores", scores, "ms", round(a.elapsed_time(b) / 5, 3))
''' % ROOT
for v in (0, 1, 2, 3, 4, 7):
    env = dict(os.environ, RQK_SCORE_DEBUG=str(v))
    out = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=300)
    print("variant", v, "|", " ; ".join(out.stdout.strip().splitlines()), out.stderr[-300:] if out.returncode else "")
