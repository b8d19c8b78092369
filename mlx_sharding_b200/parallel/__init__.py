"""Stage-to-stage transports, pipeline runtimes, fused P2P boundary and expert parallelism."""
