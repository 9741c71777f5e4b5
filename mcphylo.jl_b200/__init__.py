"""mcphylo.jl_b200 — B200-native Felsenstein likelihood + branch-length gradient behind the
PhyloDist logpdf / gradlogpdf interface of MCPhylo.jl.

Import as `mcphylo_jl_b200` (see the shim package of that name).  Host-side modules mirror
the reference's names; all device work goes through the C-ABI library built from csrc/
(include/mcphylo_b200.h).  There is no CPU fallback: evaluating without the CUDA library or
without a GPU raises.
"""
from .tree import (GeneralNode, Node, ParseNewick, newick, post_order, pre_order, get_leaves,  # noqa: F401
                   get_mother, find_by_name, find_num, number_nodes, get_branchlength_vector,
                   set_branchlength_vector, tree_length, NNI, flatten, FlatTree)
from .substitution_models import (Restriction, JC, GTR, freeK, setmatrix, freeK_equilibrium,  # noqa: F401
                                  model_derivatives, normalised_rate_matrix)
from .rates import discrete_gamma_rates, discrete_gamma_rates_dalpha, mean_boundaries, median_boundaries  # noqa: F401
from .parser import (ParseNexus, ParseCSV, datafortree, codesfortree, dense_to_codes,  # noqa: F401
                     get_alphabet, FileSyntaxError)
from .phylodist import (PhyloDist, MultiplePhyloDist, DeviceAlignment, DimensionMismatch, logpdf,  # noqa: F401
                        gradlogpdf, gradlogpdf_rates, gradlogpdf_model, multi_gradlogpdf, minimum, maximum, size, get_context,
                        set_default_device, release_device_cache)
from . import phylodist as _phylodist
globals()["__logpdf"] = getattr(_phylodist, "__logpdf")
from .prior import (exponentialBL, CompoundDirichlet, UniformBranchLength, internal_external,  # noqa: F401,E402
                    internal_logpdf, insupport, logpdfgrad)
from .synthetic import random_tree, simulate_codes  # noqa: F401,E402
from .dist import ShardedEvaluator, PipelinedEvaluator, shard_bounds, local_shard  # noqa: F401,E402
