"""CPU oracle for the CameraCalibration.correct() hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``imgprocessor_b200/`` imports this
package; only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may.  The product path has no CPU
fallback and fails loudly when ``libimgcorr.so`` is missing.

Two layers:

* ``oracle.refpath``  – a restatement of the reference's own Python chain
  (radjkarl/imgProcessor 0.2.5, ``camera/CameraCalibration.py:351-459``) that
  calls the very same third-party natives the reference calls
  (``scipy.ndimage.median_filter``, ``cv2.getOptimalNewCameraMatrix``,
  ``cv2.initUndistortRectifyMap``, ``cv2.remap``), float64 end to end.
* ``oracle.models``   – pure-numpy models of that third-party arithmetic
  (reflect median, Brown-Conrady map, OpenCV 5-bit fixed-point bilinear remap)
  plus the float32 "staged" chain the CUDA kernels are compared with bit for bit.

Third-party arithmetic that is NOT under /root/reference (un-vendored, un-pinned
in the reference's setup.py:34-41); versions the oracle was pinned against in
the build container:  scipy 1.18.1, opencv-python 4.13.0, numpy 2.3.5.

Parity pinning: the reference holds NO golden vectors / known-answer tests for
this path.  The oracle is therefore pinned against outputs
of the *unmodified reference run in the build container* under an import shim:
``tests/golden/make_golden.py`` generated ``tests/golden/*.npz`` and
``tests/test_oracle_golden.py`` checks both oracle layers against them.
"""
