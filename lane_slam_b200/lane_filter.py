"""The histogram lane filter behind the reference's interface: ``LaneFilterB200(configuration)`` mirrors
``lane_filter.LaneFilterHistogram`` (src/lane_filter/include/lane_filter/lane_filter.py:12-120: same configuration keys, same
grid, prior and Gaussian mask), but consumes the ground segments that are already on the GPU: ``process_batch`` runs
predict / update / getEstimate / getMax for every frame of the FrontEnd's last batch in one kernel (liblsf.so,
lsf_lane_filter_batch) and returns the per-frame estimates the node publishes as LanePose
(src/lane_filter/src/lane_filter_node.py:53-80)."""
import copy
import ctypes as C

import numpy as np

from . import _lib
from .frontend import check_detector_configuration  # noqa: F401  (same Configurable contract)

PARAM_NAMES = ['mean_d_0', 'mean_phi_0', 'sigma_d_0', 'sigma_phi_0', 'delta_d', 'delta_phi', 'd_max', 'd_min', 'phi_max', 'phi_min',
               'cov_v', 'linewidth_white', 'linewidth_yellow', 'lanewidth', 'min_max', 'sigma_d_mask', 'sigma_phi_mask']

# src/duckietown/config/baseline/lane_filter/lane_filter_node/default.yaml
DEFAULT_CONFIGURATION = dict(mean_d_0=0, mean_phi_0=0, sigma_d_0=0.1, sigma_phi_0=0.1, delta_d=0.02, delta_phi=0.1, d_max=0.3, d_min=-0.15,
                             phi_min=-1.5, phi_max=1.5, cov_v=0.5, linewidth_white=0.05, linewidth_yellow=0.025, lanewidth=0.23,
                             min_max=0.1, sigma_d_mask=1.0, sigma_phi_mask=2.0)


def gaussian_mask_weights(sigma, truncate=4.0):
    """Weights at distance 0 .. r of scipy.ndimage.gaussian_filter1d's kernel (scipy/ndimage/_filters.py, _gaussian_kernel1d):
    r = int(truncate * sigma + 0.5); exp(-0.5 / sigma^2 * x^2) normalised by the sum over -r .. r."""
    sd = float(sigma)
    lw = int(truncate * sd + 0.5)
    x = np.arange(-lw, lw + 1)
    phi_x = np.exp(-0.5 / (sd * sd) * x ** 2)
    phi_x = phi_x / phi_x.sum()
    return np.ascontiguousarray(phi_x[lw:]), lw


class LaneFilterConfig(C.Structure):
    _fields_ = [("nd", C.c_int32), ("nphi", C.c_int32), ("r_d", C.c_int32), ("r_phi", C.c_int32),
                ("d_min", C.c_double), ("d_max", C.c_double), ("phi_min", C.c_double), ("phi_max", C.c_double),
                ("delta_d", C.c_double), ("delta_phi", C.c_double),
                ("d_grid", C.c_void_p), ("phi_grid", C.c_void_p), ("sin_phi", C.c_void_p), ("w_d", C.c_void_p), ("w_phi", C.c_void_p),
                ("belief0", C.c_void_p)]


class LaneFilterB200(object):
    def __init__(self, configuration, front_end):
        if not isinstance(configuration, dict):
            raise ValueError('Expecting a dict, obtained %r' % (configuration,))
        configuration = copy.deepcopy(configuration)
        extra, missing = set(configuration) - set(PARAM_NAMES), set(PARAM_NAMES) - set(configuration)
        if extra or missing:            # duckietown_utils/parameters.py:15-23
            raise ValueError('Error while loading configuration.\nExtra parameters: %r\nMissing parameters: %r\n' % (extra, missing))
        for k, v in configuration.items():
            setattr(self, k, v)
        self.fe = front_end
        self._lib = front_end._lib
        # lane_filter.py:38-43
        self.d, self.phi = np.mgrid[self.d_min:self.d_max:self.delta_d, self.phi_min:self.phi_max:self.delta_phi]
        self.mean_0 = [self.mean_d_0, self.mean_phi_0]
        self.cov_0 = [[self.sigma_d_0, 0], [0, self.sigma_phi_0]]
        self.cov_mask = [self.sigma_d_mask, self.sigma_phi_mask]
        pos = np.empty(self.d.shape + (2,))
        pos[:, :, 0] = self.d
        pos[:, :, 1] = self.phi
        from scipy.stats import multivariate_normal      # the prior is the reference's own expression (:114-120)
        self.belief0 = np.ascontiguousarray(multivariate_normal(self.mean_0, self.cov_0).pdf(pos))
        wd, rd = gaussian_mask_weights(self.sigma_d_mask)
        wp, rp = gaussian_mask_weights(self.sigma_phi_mask)
        self._tables = [np.ascontiguousarray(self.d), np.ascontiguousarray(self.phi), np.ascontiguousarray(np.sin(self.phi)), wd, wp, self.belief0]
        c = LaneFilterConfig()
        c.nd, c.nphi, c.r_d, c.r_phi = self.d.shape[0], self.d.shape[1], rd, rp
        c.d_min, c.d_max, c.phi_min, c.phi_max = float(self.d_min), float(self.d_max), float(self.phi_min), float(self.phi_max)
        c.delta_d, c.delta_phi = float(self.delta_d), float(self.delta_phi)
        c.d_grid, c.phi_grid, c.sin_phi, c.w_d, c.w_phi, c.belief0 = [t.ctypes.data for t in self._tables]
        self.fe._check(self._lib.lsf_lane_filter_init(self.fe._ctx, C.byref(c)))

    def initialize(self):
        self.fe._check(self._lib.lsf_lane_filter_reset(self.fe._ctx))

    def process_batch(self, dt=None, v=None, w=None):
        """Filter step for every frame of the front end's last batch.  dt, v, w: per-frame arrays (time since the previous
        frame, linear and angular velocity; lane_filter_node.py:55) or None to skip the prediction (use_propagation: False).
        -> estimates f64 [n_frames, 3]: d, phi (getEstimate) and the belief maximum (getMax) after each frame."""
        n = self.fe._last_n
        est = np.empty((n, 3), np.float64)
        if dt is None:
            self.fe._check(self._lib.lsf_lane_filter_batch(self.fe._ctx, None, 0, est.ctypes.data))
        else:
            dvw = np.ascontiguousarray(np.stack([np.broadcast_to(np.asarray(a, np.float64), (n,)) for a in (dt, v, w)], axis=1))
            self.fe._check(self._lib.lsf_lane_filter_batch(self.fe._ctx, dvw.ctypes.data, 1, est.ctypes.data))
        return est

    @property
    def belief(self):
        out = np.empty(self.d.shape, np.float64)
        self.fe._check(self._lib.lsf_lane_filter_belief(self.fe._ctx, out.ctypes.data))
        return out

    def getEstimate(self):
        b = self.belief
        maxids = np.unravel_index(b.argmax(), b.shape)
        return [self.d_min + (maxids[0] + 0.5) * self.delta_d, self.phi_min + (maxids[1] + 0.5) * self.delta_phi]

    def getMax(self):
        return self.belief.max()
