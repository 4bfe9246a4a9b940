"""Non-canonical witnesses (tests/witness_shapes.py): the oracle on CPU, the CUDA library on the GPU."""
import pytest

from ppd_oracle_lib import parse_pre_image_dump
from witness_shapes import PAIRS


@pytest.mark.parametrize("name,weird,canonical", PAIRS, ids=[p[0] for p in PAIRS])
def test_oracle_reinsertion_makes_shape_irrelevant(oracle, name, weird, canonical):
    a = parse_pre_image_dump(oracle.compact_decode(weird))
    b = parse_pre_image_dump(oracle.compact_decode(canonical))
    assert a["state_root"] == b["state_root"]
    assert a["storage"] == b["storage"]


@pytest.mark.gpu
@pytest.mark.parametrize("name,weird,canonical", PAIRS, ids=[p[0] for p in PAIRS])
def test_gpu_general_path_matches_oracle(oracle, name, weird, canonical):
    from proof_protocol_decoder_b200 import flat
    from proof_protocol_decoder_b200.lib import Context

    ctx = Context(0)
    try:
        for w in (weird, canonical):
            got = flat.parse_pre_image_dump(ctx.compact_decode(w))
            want = parse_pre_image_dump(oracle.compact_decode(w))
            assert got["state_root"] == want["state_root"]
            assert got["storage"] == want["storage"]
            assert got["code"] == want["code"]
    finally:
        ctx.close()
