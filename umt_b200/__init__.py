"""umt_b200 — B200-native Sn sweep hot path for LLNL/UMT (Teton)."""
__version__ = "0.1.0"
