"""Drop-in alias: same import path as the reference's bayes_sim_ig/models/rff.py."""
from bayes_sim_ig_b200.models.rff import *  # noqa: F401,F403
import bayes_sim_ig_b200.models.rff as _impl

globals().update({k: v for k, v in vars(_impl).items() if not k.startswith('__')})
