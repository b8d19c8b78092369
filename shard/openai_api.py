"""``python -m shard.openai_api`` / ``mlx-sharding-api`` (reference shard/openai_api.py)."""
from mlx_sharding_b200.server.openai_api import (APIHandler, ModelProvider, convert_chat, main, run)  # noqa: F401
from mlx_sharding_b200.engine.core import stopping_criteria  # noqa: F401

if __name__ == "__main__":
    main()
