"""oracle/ — CPU restatement of the reference's detection post-processing path.

TEST INFRASTRUCTURE ONLY.  Nothing in the product package (``sar-yolo_b200/``) imports,
links or executes anything in this directory.  Allowed users: ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of
``bench.py`` — always as the checker or the timed CPU baseline, never as the shipped path.

What it restates (reference = HaoqianSong/SAR-YOLO, an Ultralytics 8.3.63 fork):
  * decode        ultralytics/nn/modules/head.py:100-131 (Detect._inference),
                  :214-249 (JDE._inference); nn/modules/block.py:62-81 (DFL);
                  utils/tal.py:366-390 (make_anchors, dist2bbox)
  * NMS glue      ultralytics/utils/ops.py:167-316 (non_max_suppression),
                  :416-433 (xywh2xyxy)
  * suppression   torchvision.ops.nms CPU kernel — THIRD PARTY, not vendored in the reference
                  (requirements.txt:2 pins torchvision==0.17.2; pyproject.toml:74 >=0.9.0).
                  Its published greedy algorithm is restated in ``nms_greedy.c``.

Parity pinning: the reference's own tests hold NO golden vector / known-answer test for this
path.  The oracle is therefore pinned against *outputs of the reference
itself run in the build container* (``oracle/ref_shim.py`` imports /root/reference live;
``tests/golden/make_golden.py`` wrote ``tests/golden/*.npz`` with it) and, wherever torchvision
is importable, against ``torchvision.ops.nms`` on CPU.  Both checks run in ``-m "not gpu"``.
"""
