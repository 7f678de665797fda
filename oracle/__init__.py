"""CPU oracle for the HFNet-SLAM per-frame front-end hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is on the product path:
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import it, and there only as the checker or as the
timed CPU arm, never as the thing shipped.

PARITY UNPINNED (network part): the reference repository ships no golden vectors,
no model weights and no automated tests, and none of
TensorRT / TensorFlow / Eigen / OpenCV-C++ exist in this image, so the reference
binary cannot be run.  The oracle is therefore a *restatement* of the reference
sources (each function cites the file:line it follows), pinned where a live
third-party implementation of the same arithmetic exists in this image:

* ``cv2.BFMatcher(NORM_L2, crossCheck=True)``  -> ``match_ref.search_by_bow``
* ``cv2.resize(INTER_LINEAR)``                 -> ``pyramid_ref.compute_pyramid``
* ``cv2.normalize(NORM_L2)``                   -> ``select_ref.l2_normalize_rows``
* ``torch.nn.functional.max_pool2d``           -> ``select_ref.simple_nms``
* ``scipy`` dense solves / finite differences  -> ``lba_ref`` Jacobians

See DESIGN.md section "Oracle" for the per-function pin status.
"""
