from lqg_b200.infer.mle import max_likelihood
from lqg_b200.infer.models import get_model_params
from lqg_b200.infer.utils import infer

__all__ = ["infer", "max_likelihood", "get_model_params"]
