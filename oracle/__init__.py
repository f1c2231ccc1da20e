"""CPU oracle for the I3D snippet-feature hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import anything from this package, and only as the checker or the timed CPU baseline.
The product package ``anomaly_detection_on_video_b200`` never imports it and has no CPU path.

Every function restates one piece of jinmang2/anomaly_detection_on_video and cites the reference
file:line it follows (paths relative to the reference root):

    preprocess.py   src/gtransforms.py:9-73,115-132, src/dataset.py:145-195 (+ Pillow's bilinear
                    resample and torchvision's Resize/TenCrop geometry, third-party, restated)
    i3res50.py      src/i3d.py:60-121,198-318 (fp32 torch functional ops == the reference arithmetic)
    segment.py      extract_features.py:77-102,159-185, src/dataset.py:121-124
    mgfn.py         src/models/mgfn/modeling_mgfn.py, src/loss/*.py

Parity pin: the reference ships no tests or golden vectors, so the pins are
outputs of the reference's own code executed in the build container on seeded inputs.
``oracle/make_golden.py`` imports ``/root/reference`` unmodified (with empty stub modules for the
absent ``pytorchvideo`` / ``decord``) and writes ``tests/golden/*.npz``; ``tests/test_oracle_*.py``
check this restatement against those fixtures (and against the live reference when it is present).
"""
