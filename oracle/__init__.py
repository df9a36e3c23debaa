"""CPU oracle of the YOLOX detection hot path — TEST INFRASTRUCTURE ONLY.

A restatement of the reference's algorithm (yhenon/pixeltable-yolox, /root/reference) used as the
checker in tests/, in __graft_entry__.smoke() and as the `cpu_baseline` / `--impl reference` arm of
bench.py. Nothing in the product package (pixeltable_yolox_b200/) imports it, and it is never the
thing that is measured as "ours" or shipped.

Pinning: every function here is checked against outputs of the reference itself (and of the
torchvision build the reference calls), generated in the build container by
tests/golden/make_golden.py and committed under tests/golden/ (see tests/test_oracle_golden.py).
The reference's own golden vectors (tests/test_detections.py:52-226) need pretrained weights from
the network and cannot be used offline.
"""
