"""C-ABI checks that need no GPU: the library loads, exports every symbol include/graspa_b200.h declares,
fails loudly without a device, and the product never touches oracle/."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    with open(os.path.join(ROOT, "include", "graspa_b200.h")) as f:
        src = f.read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(gb_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from graspa_b200 import engine
    engine.build()
    lib = engine.load_library()
    syms = header_symbols()
    assert len(syms) >= 40
    missing = [s for s in syms if not hasattr(lib, s)]
    assert not missing, missing
    assert sorted(engine.DECLARED_SYMBOLS) == syms
    import re
    want = int(re.search(r"#define GB_ABI_VERSION (\d+)", open(os.path.join(ROOT, "include", "graspa_b200.h")).read()).group(1))
    assert lib.gb_abi_version() == want == 6      # 3: block pockets, identity-swap commit; 4: CB/CFC lambda change; 5: NPT volume move; 6: gb_widom_inputs.sums_device


def test_engine_creation_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from graspa_b200 import engine
    with pytest.raises(engine.EngineError, match="no CUDA device"):
        engine.Engine()


def test_product_does_not_reference_oracle():
    bad = []
    for base, _, files in os.walk(os.path.join(ROOT, "graspa_b200")):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h", ".inc", ".cpp", ".hpp")) or fn == "Makefile":
                with open(os.path.join(base, fn), errors="ignore") as f:
                    t = f.read()
                if re.search(r"(import\s+oracle|from\s+oracle|liboracle|graspa_oracle|oracle/)", t):
                    if not re.search(r"nothing .* oracle|never .* oracle|includes, links or calls oracle", t):
                        bad.append(fn)
    assert not bad, bad


def test_integration_adapter_snippet_compiles_against_the_header(tmp_path):
    """The reference-side adapter shown in INTEGRATION.md (section 1) must stay in step with include/graspa_b200.h: it is cut
    out of the document and compiled (syntax + types) against stand-ins of the reference structs it reads."""
    import shutil
    import subprocess
    gxx = shutil.which("g++")
    if not gxx:
        pytest.skip("no g++")
    doc = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    start = doc.index("```cpp") + 6
    code = doc[start:doc.index("```", start)].replace('#include "data_struct.h"', '#include "ref_stub.h"')
    src = tmp_path / "adapter.cpp"
    src.write_text(code + "\nint main() { return 0; }\n")
    r = subprocess.run([gxx, "-std=c++17", "-fsyntax-only", "-I", os.path.join(ROOT, "include"), "-I", os.path.join(ROOT, "tests", "support"), str(src)],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
