from flash_cosine_sim_attention_b200.flash_cosine_sim_attention import *  # noqa: F401,F403
from flash_cosine_sim_attention_b200.flash_cosine_sim_attention import backward, debug, forward  # noqa: F401
