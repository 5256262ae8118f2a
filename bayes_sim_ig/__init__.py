"""Drop-in alias package: the reference's import paths (``bayes_sim_ig.bayes_sim``,
``bayes_sim_ig.models.mdnn`` ...) backed by ``bayes_sim_ig_b200``."""
