"""CPU oracle for the miso / torchvision detection post-processing hot path.

TEST INFRASTRUCTURE ONLY — the product (miso_b200/) never imports this package. Only
tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may,
and only as the checker or the timed CPU baseline.

What is restated, and from where (citations: `ref:` = microfossil/particle-object-detection,
`tv:` = torchvision 0.26.0 Python layer, `tv-csrc:` = torchvision native sources, which are
not on disk — their algorithm is the one pinned in SURVEY.md Appendix B):

  oracle/c/oracle_kernels.c   greedy NMS and RoIAlign forward (the two native CPU kernels)
  oracle/detection.py         anchors, BoxCoder.decode, clip, remove_small, per-level top-k,
                              batched_nms (both strategies), filter_proposals, LevelMapper,
                              MultiScaleRoIAlign, postprocess_detections, resize_boxes, box_convert
  oracle/miso_path.py         miso's score filter, xyxy->xywh annotation, coords_int and crop slice
  oracle/mosaic.py            tile grid, rank partition and the seam-NMS composition (config 5;
                              no reference counterpart — semantics defined in SURVEY.md §8(e))

Pinning: the reference repo has no tests, fixtures or golden vectors, so the
oracle is pinned against outputs of the reference arithmetic itself: the importable
torchvision 0.26.0 CPU ops. tests/gen_golden.py generated tests/golden/*.npz from those ops
(script committed), and tests/test_oracle_pin.py re-checks live against torchvision wherever
it is importable.
"""
