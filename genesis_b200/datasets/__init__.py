"""Synthetic stand-ins for the reference's datasets/*_config.py (same Forge `load(cfg)` contract)."""
