"""Drop-in mirror of the reference's `src` package surface for the DB1 hot path (src.model, src.mpu, the input
dataclasses and the tokenizers the model imports). Put `bdm-db1_b200/` first on sys.path and the reference's training /
evaluation loops import these instead of their own."""
