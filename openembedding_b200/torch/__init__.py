"""``import openembedding_b200.torch as embed`` -- mirror of ``import openembedding.tensorflow as embed``."""
from ..api import *  # noqa: F401,F403
from ..api import (Adadelta, Adagrad, Adam, Adamax, Embedding, Ftrl, FtrlDistributed, Model, ModelCheckpoint, Nadam,
                   RMSprop, SGD,
                   Variable, distributed_model, distributed_optimizer, distributed_variable, load_server_model,
                   persist_server_model, pulling, restore_server_model, save_as_original_model, save_server_model,
                   should_persist_server_model)
from .. import flags, Master, Server  # noqa: F401
