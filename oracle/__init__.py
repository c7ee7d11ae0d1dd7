"""CPU oracle of the PIGEON inference hot path — TEST INFRASTRUCTURE ONLY.

A plain PyTorch-fp32 / numpy restatement of the reference's algorithm for the path
(CLIP ViT forward -> token mean -> geocell head -> ProtoRefiner), each function citing the
reference file:line it follows.  Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s
`cpu_baseline` / `--impl reference` legs may import this package, and only as the checker or the
timed CPU baseline — never from `pigeon_b200/` (the product path has no CPU fallback).

Parity status: PINNED against the reference's own modules.  The reference ships no tests or golden
vectors, so `oracle/make_golden.py` imports the UNMODIFIED reference
`models/super_guessr.py`, `models/proto_refiner.py`, `preprocessing/geo_utils.py`,
`preprocessing/utils.py` from /root/reference behind `oracle/reference_shim.py`, runs them on seeded
inputs and commits the outputs under `tests/golden/`; `tests/test_oracle_golden.py` checks this
restatement against those fixtures.  The ViT arithmetic itself lives in third-party
`transformers` (pinned 4.23.1 by the reference's env.yml:60; 5.5.0 installed here, eager attention) —
its `CLIPVisionModel` is what the reference calls at models/clip_embedder.py:63 and
models/super_guessr.py:395, and the golden vectors are produced through exactly that call.
"""
