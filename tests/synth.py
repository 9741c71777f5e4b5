from mcphylo_jl_b200.synthetic import random_tree, simulate_codes  # noqa: F401
