"""Generate tests/golden/odometry.npz by running the REFERENCE's own OdometryNode (src/odometry/src/odometry.py, loaded
unmodified from /root/reference) on a synthetic wheels-command log.  Authoring container only.

Shims: rospy (Time.now, Subscriber, Publisher), tf (TransformBroadcaster, transformations.quaternion_from_euler),
message classes -- only what the node touches.  Stored: the command log (stamp nsecs, vel_left, vel_right) and, after
every command, the node's (pos, theta) and whether it broadcast a transform (0 < dt < 0.3)."""
import importlib.util
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/src/odometry/src/odometry.py"


class _Obj(object):
    def __init__(self, **kw):
        self.__dict__.update(kw)


def load_reference_node():
    now = _Obj(nsecs=0)
    rospy = types.ModuleType("rospy")
    rospy.Time = _Obj(now=lambda: now)
    rospy.Subscriber = lambda *a, **k: None
    rospy.Publisher = lambda *a, **k: _Obj(publish=lambda m: None)
    sys.modules["rospy"] = rospy
    for name, attrs in (("std_msgs", []), ("std_msgs.msg", ["String"]), ("duckietown_msgs", []), ("duckietown_msgs.msg", ["WheelsCmdStamped"]),
                        ("visualization_msgs", []), ("visualization_msgs.msg", ["Marker"]), ("geometry_msgs", []), ("geometry_msgs.msg", ["Point"])):
        m = types.ModuleType(name)
        for a in attrs:
            setattr(m, a, type(a, (object,), {}))
        sys.modules[name] = m

    class Marker(object):
        LINE_STRIP, ADD = 4, 0

        def __init__(self):
            self.header = _Obj(stamp=None, frame_id="")
            self.scale = _Obj(x=0, y=0, z=0); self.color = _Obj(a=0, r=0, g=0, b=0)
            self.points = []
    sys.modules["visualization_msgs.msg"].Marker = Marker
    sys.modules["geometry_msgs.msg"].Point = lambda: _Obj(x=0.0, y=0.0, z=0.0)
    tf = types.ModuleType("tf")
    sent = []
    tf.TransformBroadcaster = lambda: _Obj(sendTransform=lambda *a: sent.append(a))
    tf.transformations = _Obj(quaternion_from_euler=lambda r, p, y: (0.0, 0.0, np.sin(y / 2), np.cos(y / 2)))
    sys.modules["tf"] = tf
    spec = importlib.util.spec_from_file_location("ref_odometry", REF)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    mod.br = tf.TransformBroadcaster()
    return mod, sent


def main():
    mod, sent = load_reference_node()
    node = mod.OdometryNode()
    rng = np.random.default_rng(11)
    n = 600
    # stamps: ~10 Hz in the nsecs field, with wrap-arounds (negative dt), stalls (dt = 0) and gaps (dt >= 0.3) like a real log
    dts = rng.uniform(0.02, 0.2, n)
    dts[rng.integers(0, n, 12)] = 0.0
    dts[rng.integers(0, n, 12)] = rng.uniform(0.3, 0.9, 12)
    t = np.cumsum(dts) % 1.0                       # the reference only reads stamp.nsecs
    nsecs = np.floor(t * 1e9)
    vl = np.round(rng.uniform(-0.2, 0.6, n), 3)
    vr = np.round(vl + rng.choice([0.0, 0.0, 0.05, -0.08, 0.3], n) * rng.uniform(0, 1, n).round(2), 3)
    out = np.zeros((n, 4), np.float64)
    for i in range(n):
        before = len(sent)
        node.getPose(_Obj(header=_Obj(stamp=_Obj(nsecs=float(nsecs[i]))), vel_left=float(vl[i]), vel_right=float(vr[i])))
        out[i] = [node.pos[0], node.pos[1], node.theta, len(sent) - before]
    path = os.path.join(HERE, "odometry.npz")
    np.savez_compressed(path, nsecs=nsecs, vel_left=vl, vel_right=vr, poses=out[:, :3], advanced=out[:, 3].astype(np.int32))
    print("wrote", path, "advanced", int(out[:, 3].sum()), "of", n, "straight steps", int((vl == vr).sum()))


if __name__ == "__main__":
    main()
