"""Minimal stand-in for mmcv-full 1.4.8 so the unmodified reference imports offline (test infrastructure)."""
