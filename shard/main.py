"""``python -m shard.main`` / ``mlx-sharding-server`` (reference shard/main.py)."""
from mlx_sharding_b200.server.shard_server import main, serve  # noqa: F401

if __name__ == "__main__":
    main()
