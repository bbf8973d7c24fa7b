from bayesformer_b200.configs.constants import *  # noqa: F401,F403
from bayesformer_b200.configs.constants import EOS_TOKEN, PAD_TOKEN, RANDOM_SEED  # noqa: F401
