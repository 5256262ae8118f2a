"""Drop-in alias: same import path as the reference's bayes_sim_ig/utils/pdf.py."""
from bayes_sim_ig_b200.utils.pdf import *  # noqa: F401,F403
import bayes_sim_ig_b200.utils.pdf as _impl

globals().update({k: v for k, v in vars(_impl).items() if not k.startswith('__')})
