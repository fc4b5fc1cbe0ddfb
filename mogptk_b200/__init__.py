"""mogptk_b200 -- B200-native exact multi-output GP engine behind the mogptk.gpr interface.

Hot path: MOSM / SM / CONV Gram build, blocked Cholesky, log-marginal
likelihood with analytic gradient, posterior mean/variance -- hand-written sm_100a CUDA in
libmogp_b200.so, called through the C ABI of include/mogp_b200.h.  No CPU fallback.
"""
from . import _cabi, engine, gpr, init, synth    # noqa: F401
from .engine import Engine, NotPositiveDefiniteError  # noqa: F401
from .gpr import (CholeskyException, CrossSpectralKernel, Exact, GaussianConvolutionProcessKernel,  # noqa: F401
                  GaussianLikelihood, IndependentMultiOutputKernel, LinearModelOfCoregionalizationKernel, MixtureKernel,
                  MultiOutputHarmonizableSpectralKernel, MultiOutputSpectralMixtureKernel, Parameter, SpectralKernel,
                  SpectralMixtureKernel, UncoupledMultiOutputSpectralKernel)
from .inference import B200Exact                 # noqa: F401
from .train import fit_adam, install, uninstall  # noqa: F401

__version__ = "0.1.0"
