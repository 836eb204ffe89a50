"""Synthetic user workloads named by BASELINE.json's configs (what would run inside the reference's
training container).  Benchmark inputs, not product code."""
