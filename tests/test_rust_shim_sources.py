"""The Rust shim of INTEGRATION.md (integration/rust/: source only, no rustc in this image) against the C headers it binds.

What can be checked without a Rust toolchain: every `extern "C"` declaration names a function include/ppd_b200.h
declares, with the same number of arguments; the format constants equal include/ppd_flat.h's; every status of
include/ppd_status.h is handled by status.rs; INTEGRATION.md shows the files as they are."""
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
RUST = os.path.join(ROOT, "integration", "rust")


def _read(*p):
    with open(os.path.join(*p)) as f:
        return f.read()


def _strip_c_comments(s):
    out, i = [], 0
    while i < len(s):
        j = s.find("/*", i)
        if j < 0:
            out.append(s[i:])
            break
        out.append(s[i:j])
        k = s.find("*/", j + 2)
        i = len(s) if k < 0 else k + 2
    return "".join(out)


def _c_prototypes():
    """name -> argument count of every function include/ppd_b200.h declares."""
    h = _strip_c_comments(_read(ROOT, "include", "ppd_b200.h"))
    protos = {}
    for stmt in h.split(";"):
        stmt = " ".join(stmt.split())
        m = re.search(r"\b(ppd_[a-z0-9_]+)\s*\(([^()]*)\)$", stmt)
        if not m or "typedef" in stmt:
            continue
        args = m.group(2).strip()
        protos[m.group(1)] = 0 if args in ("", "void") else args.count(",") + 1
    return protos


def _rust_externs():
    src = _read(RUST, "src", "gpu_ffi.rs")
    block = src[src.index('extern "C" {') :]
    block = block[: block.index("\n}\n")]
    fns = {}
    for m in re.finditer(r"pub fn (ppd_[a-z0-9_]+)\(([^)]*)\)", block):
        args = m.group(2).strip()
        fns[m.group(1)] = 0 if not args else len([a for a in args.split(",") if a.strip()])
    return fns


def test_every_extern_declaration_matches_the_header():
    protos, externs = _c_prototypes(), _rust_externs()
    assert len(externs) >= 10
    for name, n_args in externs.items():
        assert name in protos, f"gpu_ffi.rs binds {name}, which include/ppd_b200.h does not declare"
        assert protos[name] == n_args, f"{name}: {n_args} arguments in gpu_ffi.rs, {protos[name]} in the header"
    # the entry points the shim's into_txn_proof_gen_ir and a batching caller need
    for name in ("ppd_ctx_create", "ppd_ctx_destroy", "ppd_last_error", "ppd_free", "ppd_block_decode", "ppd_blocks_decode_batch", "ppd_blocks_decode_stream"):
        assert name in externs


def test_format_constants_equal_the_headers():
    h = _read(ROOT, "include", "ppd_flat.h")
    defines = {m.group(1): int(m.group(2), 16) for m in re.finditer(r"#define (PPD_[A-Z_]+) (0x[0-9a-fA-F]+)", h)}
    rs = _read(RUST, "src", "b200", "flat.rs")
    consts = {m.group(1): int(m.group(2).replace("_", ""), 16) for m in re.finditer(r"const ([A-Z_]+): u(?:8|32) = (0x[0-9a-fA-F_]+);", rs)}
    assert consts["FLAT_BLOCK_MAGIC"] == defines["PPD_FLAT_BLOCK_MAGIC"]
    assert consts["IR_DUMP_MAGIC"] == defines["PPD_IR_DUMP_MAGIC"]
    flags = [k for k in defines if k.startswith("PPD_TR_")]
    assert len(flags) == 7
    for k in flags:
        assert consts[k[len("PPD_") :]] == defines[k], k
    # and the Python twin of the same encoder uses the same numbers
    from proof_protocol_decoder_b200 import flat

    assert flat.IR_DUMP_MAGIC == defines["PPD_IR_DUMP_MAGIC"]


def test_every_status_has_an_arm_in_status_rs():
    h = _read(ROOT, "include", "ppd_status.h")
    codes = {int(m.group(2)): m.group(1) for m in re.finditer(r"(PPD_[A-Z0-9_]+) = (\d+)", h)}
    rs = _read(RUST, "src", "b200", "status.rs")
    body = rs[rs.index("pub fn status_to_error") :]
    handled = set()
    for m in re.finditer(r"^\s+((?:\d+(?:\.\.=\d+)?)(?: \| \d+)*) =>", body, re.M):
        for part in m.group(1).split(" | "):
            if "..=" in part:
                a, b = part.split("..=")
                handled.update(range(int(a), int(b) + 1))
            else:
                handled.add(int(part))
    for code, name in codes.items():
        if code == 0:
            continue
        if name == "PPD_ERR_NON_EXISTENT_TRIE_ENTRY":
            assert code not in handled  # declared by the reference, constructed nowhere (decoding.rs:40): never returned
            continue
        assert code in handled, f"status.rs has no arm for {name} = {code}"
    # the variant names of CompactParsingError, in the header's (= the enum's) order
    names = re.search(r"COMPACT_VARIANTS: \[&str; 11\] = \[(.*?)\];", rs, re.S).group(1)
    names = [n.strip().strip('"') for n in names.split(",") if n.strip()]
    assert len(names) == 11
    for i, n in enumerate(names, start=1):
        want = codes[i][len("PPD_ERR_") :].replace("_", "").lower()
        assert n.lower().startswith(want[:12]), (i, n, codes[i])


def test_status_rs_reads_the_payload_words_err_detail_h_writes():
    rs = _read(RUST, "src", "b200", "status.rs")
    hd = _read(ROOT, "proof_protocol_decoder_b200", "csrc", "err_detail.h")
    for word in ("bytes", "hashed_addr", "trie_type", "addr", "amount"):
        assert f'"{word}"' in rs and f"{word}=" in hd, word
    for variant in ("State", "Storage", "Receipt", "Txn"):
        assert f'"{variant}"' in hd and variant in rs


def test_integration_md_shows_the_files_as_they_are():
    res = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "gen_integration_md.py"), "--check"], timeout=60)
    assert res.returncode == 0, "INTEGRATION.md is stale: run python tools/gen_integration_md.py"
