from flash_cosine_sim_attention_b200.transformer import Attention, CosineSimCausalTransformer, FeedForward  # noqa: F401
