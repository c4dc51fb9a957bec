from flash_cosine_sim_attention_b200.benchmark import benchmark  # noqa: F401
