"""panfeed_b200 — B200-native build of panfeed's per-gene-cluster k-mer
streaming hot path."""
__version__ = "0.1.0"
