"""CPU oracle for the 2D ICP scan-matching hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product:
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import it, and there only as the checker or the
timed CPU baseline.  The shipped path (``icp_slam-yolo_b200``) never imports
this package and fails loudly when its CUDA library is missing.

Parity status: PINNED.  ``oracle.icp_oracle`` is a restatement of
``/root/reference/labels_segmentation/icp.py:5-53`` and
``/root/reference/duc/ICP_LIDAR/process.py:9-52``; ``tests/golden/make_golden.py``
imports the unmodified reference in the build container and records its outputs
(circle demo known-answer test, Scan_data_1 spot pairs) as fixtures under
``tests/golden/``; ``tests/test_oracle.py`` checks the restatement against them
bit for bit (and against the live reference whenever ``/root/reference`` is
present).  The gated / initial-pose extensions follow Open3D's call shape and
are parity-UNPINNED (Open3D is not vendored by the reference).
"""
