"""Entry points: shard server (``mlx-sharding-server``), OpenAI-compatible API (``mlx-sharding-api``)."""
