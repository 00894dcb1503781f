"""scene/beta_model.py:13-17 imports three `_torch_impl` names and never calls them.  They resolve here to the CUDA
operators (same signatures and results) so that the import succeeds without shipping a second implementation."""
from ubs_b200.dropin import cond_mean_convariance_opacity as _cond_mean_convariance_opacity  # noqa: F401
from ubs_b200.dropin import l_triangle_to_rotmat as _l_triangle_to_rotmat  # noqa: F401
from ubs_b200.dropin import rot_scale_l_triangle_to_covar as _rot_scale_l_triangle_to_covar  # noqa: F401
