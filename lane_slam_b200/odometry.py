"""Diff-drive odometry of the reference (src/odometry/src/odometry.py:43-120), minus ROS: same state (pos, theta, last_t, dt),
same update rule; the arithmetic is lsf_odometry_step in liblsf.so (host double precision, like the reference's Python).
``poses()`` gives the [n, 3] array lsf_map_append takes."""
import ctypes as C

import numpy as np

from . import _lib


def nccl_unique_id():
    """128-byte NCCL unique id (rank 0 creates it; broadcast it to the other ranks)."""
    buf = (C.c_char * 128)()
    rc = _lib.load().lsf_nccl_unique_id(buf)
    if rc != 0:
        raise _lib.LsfError(rc, _lib.load().lsf_last_error(None).decode())
    return bytes(buf)


class Odometry(object):
    """OdometryNode without the ROS plumbing: getPose(stamp_nsecs, vel_left, vel_right) -> True when the pose advanced."""

    def __init__(self, stamp_nsecs=0.0):
        self._lib = _lib.load()
        self._st = _lib.LsfOdometry()
        self._lib.lsf_odometry_init(C.byref(self._st), float(stamp_nsecs))
        self.trajectory = []                       # what showTraj() appends to the Marker (odometry.py:100-106)

    @property
    def pos(self):
        return [self._st.x, self._st.y]

    @property
    def theta(self):
        return self._st.theta

    @property
    def dt(self):
        return self._st.dt

    def getPose(self, stamp_nsecs, vel_left, vel_right):
        rc = self._lib.lsf_odometry_step(C.byref(self._st), float(stamp_nsecs), float(vel_left), float(vel_right))
        if rc < 0:
            raise _lib.LsfError(rc, "lsf_odometry_step")
        if rc == 1:
            self.trajectory.append((self._st.x, self._st.y))
        return rc == 1

    def pose(self):
        return (self._st.x, self._st.y, self._st.theta)


def integrate(stamps_nsecs, vel_left, vel_right, stamp0_nsecs=0.0):
    """Poses [n, 3] after each wheels command (the pose a frame taken at that time is placed with)."""
    od = Odometry(stamp0_nsecs)
    out = np.empty((len(stamps_nsecs), 3), np.float64)
    for i, (t, vl, vr) in enumerate(zip(stamps_nsecs, vel_left, vel_right)):
        od.getPose(t, vl, vr)
        out[i] = od.pose()
    return out
