"""``B200Exact`` -- the plug-in for the reference's ``inference=`` builder seam.

The reference constructs its GP model with ``self.gpr = inference._build(kernel, x, y, y_err, mean)``
(mogptk/model.py:231; stock builder ``mogptk.Exact`` at mogptk/model.py:76-100).  This class has
the same constructor arguments and the same ``_build`` signature but returns
``mogptk_b200.gpr.Exact``, whose O(N^2)/O(N^3) work runs on the B200 engine:

    import mogptk, mogptk_b200
    model = mogptk.MOSM(dataset, Q=5, inference=mogptk_b200.B200Exact())
    model.train(method='Adam', iters=500, lr=0.1)
    model.predict()
"""
import sys

from . import gpr


class B200Exact:
    """Exact inference on the B200 engine.

    Args:
        variance (float): Variance of the Gaussian likelihood (default 1.0 per channel).
        data_variance: Fixed per-point variances added to the diagonal.
        jitter (float): Relative jitter added before the Cholesky.
        engine: optional ``mogptk_b200.engine.Engine`` to run on (default: one per device).
    """

    def __init__(self, variance=None, data_variance=None, jitter=1e-8, engine=None):
        self.variance = variance
        self.data_variance = data_variance
        self.jitter = jitter
        self.engine = engine

    def _build(self, kernel, x, y, y_err=None, mean=None):
        variance = self.variance
        if variance is None:                                   # mogptk/model.py:90-95
            variance = [1.0] * kernel.output_dims if kernel.output_dims is not None else 1.0
        data_variance = self.data_variance
        if data_variance is None and y_err is not None:        # mogptk/model.py:96-98
            data_variance = y_err ** 2
        like_cls, chol_exc = None, None
        ref = sys.modules.get("mogptk")
        if ref is not None and type(kernel).__module__.startswith("mogptk."):
            # driven by the reference: keep its likelihood / exception types and follow its device
            like_cls = ref.gpr.GaussianLikelihood
            chol_exc = ref.gpr.CholeskyException
            gpr.config.device = ref.gpr.config.device
            gpr.config.dtype = ref.gpr.config.dtype
        return gpr.Exact(kernel, x, y, variance=variance, data_variance=data_variance, jitter=self.jitter,
                         mean=mean, engine=self.engine, likelihood_cls=like_cls, cholesky_exception=chol_exc)
