"""Checkpoint IO, quantisation helpers, loader, timing utilities."""
