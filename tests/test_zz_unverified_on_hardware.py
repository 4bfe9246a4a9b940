"""GPU tests written after round 2's GPU minutes were spent: they have NEVER run on hardware.

They are skipped unless PPD_RUN_UNVERIFIED=1 — an input shape the kernels have not seen is not something to meet for the
first time inside the suite that certifies the round — and the file sorts last.  What each of them checks has a CPU
counterpart that did run (named in the test).  First thing to do with a GPU: PPD_RUN_UNVERIFIED=1 pytest -m gpu on this file."""
import os

import numpy as np
import pytest

from ppd_oracle_lib import OracleError

pytestmark = [
    pytest.mark.gpu,
    pytest.mark.skipif(os.environ.get("PPD_RUN_UNVERIFIED") != "1", reason="never run on hardware; set PPD_RUN_UNVERIFIED=1"),
]


@pytest.fixture(scope="module")
def ctx():
    from proof_protocol_decoder_b200.lib import Context

    c = Context(0)
    yield c
    c.close()


def test_block_error_payloads_match_oracle(ctx, oracle):
    """ppd_last_error carries the payload of the TraceParsingError variants as the oracle spells it (csrc/err_detail.h).
    CPU counterpart: test_txn_core_cpu.py::test_host_path_error_payloads_equal_the_oracles (the host path reports every
    block error, and the harness runs exactly that code)."""
    from proof_protocol_decoder_b200 import PpdError
    from test_txn_core_cpu import _error_blocks

    for name, code, flat_block in _error_blocks():
        with pytest.raises(OracleError) as eo:
            oracle.block_decode(flat_block)
        with pytest.raises(PpdError) as eg:
            ctx.block_decode(flat_block)
        assert eg.value.code == eo.value.code == code and eg.value.msg == eo.value.msg, name
        assert bool(eg.value.payload()) == (code in (24, 25)), name


@pytest.mark.parametrize("host_txn", [False, True])
def test_witness_like_storage_tries(ctx, oracle, host_txn, monkeypatch):
    """Storage tries with hashed siblings next to every touched path (what a mainnet witness carries; the generator
    witnesses them in full), most writes deletes: the product's IR bytes are the oracle's, through both loops.
    CPU counterpart: test_txn_core_cpu.py::test_device_txn_loop_on_witness_like_storage_tries."""
    from proof_protocol_decoder_b200 import synth
    from test_txn_core_cpu import _witness_like_storage

    if host_txn:
        monkeypatch.setenv("PPD_HOST_TXN", "1")
    else:
        monkeypatch.delenv("PPD_HOST_TXN", raising=False)
    rng = np.random.default_rng(1)
    for i in range(8):
        b = synth.gen_block(900 + i, n_accounts=40, n_txns=int(rng.integers(2, 8)), inline_code_frac=0.0, contract_frac=0.7, slots_lo=4, slots_hi=60,
                            slot_reads=(0, 4), slot_writes=(0, 6), zero_write_frac=0.6, accounts_per_txn=(2, 6))
        fb, n_hashed = _witness_like_storage(oracle, b, rng)
        assert n_hashed > 0
        assert ctx.block_decode(fb) == oracle.block_decode(fb), f"block {i}"
