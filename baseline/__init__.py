"""Baselines timed beside the B200 path (never imported by the product package)."""
