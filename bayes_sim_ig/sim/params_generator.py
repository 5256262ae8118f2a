"""Alias of bayes_sim_ig_b200.sim.params_generator (drop-in import path)."""
from bayes_sim_ig_b200.sim.params_generator import *  # noqa: F401,F403
from bayes_sim_ig_b200.sim.params_generator import ParamsSampler  # noqa: F401
