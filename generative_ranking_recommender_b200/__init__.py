"""B200-native (sm_100a) implementation of the semantic-ID hot path of
zeehu/generative_ranking_recommender: hierarchical residual-quantisation balanced K-Means.

Public surface = the reference's own:
    from generative_ranking_recommender_b200.hierarchical_rq_kmeans import (
        HierarchicalRQKMeans, HierarchicalRQKMeansConfig, CheckpointManager)
    from generative_ranking_recommender_b200.balancekmeans import KMeans, pairwise_distance_full

Importing the package never touches the GPU; the first compute call loads librqk_sm100a.so and
raises if it (or an sm_100 device) is missing - there is no CPU fallback."""

__version__ = "0.1.0"

from .hierarchical_rq_kmeans import (  # noqa: E402,F401
    CheckpointManager,
    HierarchicalRQKMeans,
    HierarchicalRQKMeansConfig,
    hierarchicalRqClusterParams,
)
