"""Generates tests/golden/c_api_symbols.txt: every `sk_*` function declared by the reference's C ABI headers
(/root/reference/cpp/include/c_api/*.h, the interface rust/sasktran2-sys binds).  The reference tree does not exist on
the GPU box, so the list is committed; tests/test_host_logic.py dlsym()s each name in libsasktran2_b200.so.

    python tests/golden/make_c_api_symbols.py [/root/reference]
"""
import re
import sys
from pathlib import Path

ref = Path(sys.argv[1] if len(sys.argv) > 1 else "/root/reference")
names = set()
for h in sorted((ref / "cpp" / "include" / "c_api").glob("*.h")):
    text = re.sub(r"/\*.*?\*/", "", h.read_text(), flags=re.S)
    text = re.sub(r"//[^\n]*", "", text)
    names.update(re.findall(r"\b(sk_[a-z0-9_]+)\s*\(", text))
out = Path(__file__).resolve().parent / "c_api_symbols.txt"
out.write_text("\n".join(sorted(names)) + "\n")
print(f"{len(names)} symbols -> {out}")
