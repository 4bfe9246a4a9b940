"""Rewrites the Rust shim section of INTEGRATION.md from the files under integration/rust/: the fenced rust block under
each heading that names a shim file.  `python tools/gen_integration_md.py --check` exits 1 when the section is stale."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FILES = {
    "protocol_decoder/build.rs": "integration/rust/build.rs",
    "protocol_decoder/src/gpu_ffi.rs": "integration/rust/src/gpu_ffi.rs",
    "protocol_decoder/src/b200/mod.rs": "integration/rust/src/b200/mod.rs",
    "protocol_decoder/src/b200/flat.rs": "integration/rust/src/b200/flat.rs",
    "protocol_decoder/src/b200/status.rs": "integration/rust/src/b200/status.rs",
}
OPEN, CLOSE = "```rust\n", "\n```\n"


def regenerate(md: str) -> str:
    for name, path in FILES.items():
        src = open(os.path.join(ROOT, path)).read().rstrip("\n")
        h = md.index("### `" + name + "`\n")
        a = md.index(OPEN, h) + len(OPEN)
        b = md.index(CLOSE, a)
        md = md[:a] + src + md[b:]
    return md


if __name__ == "__main__":
    p = os.path.join(ROOT, "INTEGRATION.md")
    old = open(p).read()
    new = regenerate(old)
    if "--check" in sys.argv:
        sys.exit(0 if new == old else 1)
    open(p, "w").write(new)
