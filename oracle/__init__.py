"""CPU ORACLE -- test infrastructure, not product code.

Restates the reference's per-frame line front end (mandanasmi/lane-slam) on the CPU:

* ``reference_glue``  -- the reference's Python glue (line_detector_lsd.py, the three ROS
  nodes' per-frame arithmetic) restated for OpenCV 4 and calling the REAL third-party
  library ``cv2`` 4.13 that the reference calls.  This is the ground truth.
* ``cmodel``          -- ctypes bindings of ``csrc/lane_oracle.c``: plain-C, cv2-free models
  of every primitive (HSV, Canny, LSD, undistort, LBD, Hamming kNN) that give stage-level
  goldens for each CUDA kernel.  Pinned against cv2 / the imported reference in tests/.
* ``synth``           -- the synthetic Duckietown-style frame generator.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / ``--impl reference`` legs
may import this package.  ``lane_slam_b200`` never does.
"""
