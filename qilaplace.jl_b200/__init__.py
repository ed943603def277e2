"""qilaplace_b200 -- B200-native (sm_100a) hot path of QILaplace.jl behind the reference's API.

The package directory is named ``qilaplace.jl_b200`` (not a legal module name); import it through
the repository-root shim ``qilaplace_b200``.
"""
from ._lib import (ArgumentError, DomainError, ErrorException, CudaError, UnsupportedError,
                   LIB_PATH, declared_symbols, load as load_library)
from .api import *  # noqa: F401,F403
from .api import (Context, default_context, SignalMPS, ZTMPS, SingleSiteMPO, PairedSiteMPO,
                  coefficient, coefficients, apply, generate_signal, signal_mps, signal_ztmps,
                  canonicalize, canonicalize_, compress, compress_, norm, mps_to_vector,
                  build_qft_mpo, build_dt_mpo, build_zt_mpo, qr, svd_trunc, rsvd,
                  signal_mps_dev, signal_mps_batch_dev, ztmps_from_mps, coefficients_dev,
                  coefficient_grid, coefficient_grid_dev, pole_scan, pole_scan_modes, apply_batch,
                  coefficient_grid_argmax, coefficients_argmax, sum_sites, laplace_coefficients, z_from_kl, kl_bits,
                  pole_scan_argmax, pole_scan_list_argmax, pole_scan_driver, save, load_mps, load_mpo, apply_zipup, Uploader)
