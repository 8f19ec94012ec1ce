"""CPU tests of the C-ABI boundary: the library loads, exports every declared symbol, and refuses
to run without a GPU (no CPU fallback)."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "gan_b200.h")).read()
    return sorted(set(re.findall(r"GAN_API\s+[\w\s\*]+?\b(gan_\w+)\s*\(", src)))


def test_library_exports_every_declared_symbol(lib):
    syms = _declared_symbols()
    assert len(syms) >= 40
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/gan_b200.h but not exported"


def test_bindings_cover_the_header():
    from gan_b200 import _ffi
    assert sorted(_ffi.SIGNATURES.keys()) == _declared_symbols()


def test_no_cpu_fallback(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    h = C.c_void_p()
    rc = lib.gan_ctx_create(0, 1, C.c_uint64(0), C.byref(h))
    assert rc == -3 and b"no CPU fallback" in lib.gan_last_error()
    from gan_b200 import Pix2Pix, _ffi
    with pytest.raises(_ffi.GanError):
        Pix2Pix({"img_size": 256, "channels": "3"})


def test_product_code_never_imports_the_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "gan_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in txt and "from oracle" not in txt, f


def test_image_xform_struct_matches_the_header():
    """ctypes mirror of gan_image_xform: same field names, order and size as include/gan_b200.h."""
    import re
    import ctypes as C
    from gan_b200 import _ffi
    hdr = open(os.path.join(ROOT, "include", "gan_b200.h")).read()
    body = re.search(r"typedef struct \{(.*?)\} gan_image_xform;", hdr, re.S).group(1)
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    names = [n.strip() for decl in re.findall(r"int ([^;]+);", body) for n in decl.split(",")]
    assert names == [f[0] for f in _ffi.ImageXform._fields_]
    assert C.sizeof(_ffi.ImageXform) == 4 * len(names)


def test_built_library_carries_the_blackwell_instructions():
    """The hot path must stay on tcgen05 / TMEM / TMA: the SASS of the built library holds the CTA-pair MMA
    (tcgen05.mma.cta_group::2), TMA tensor loads incl. the 3-D image boxes of the first-layer kernels, TMEM loads, and no
    legacy HMMA (mma.sync) instruction anywhere.  Skipped when cuobjdump is not installed."""
    import shutil
    import subprocess
    if shutil.which("cuobjdump") is None:
        pytest.skip("cuobjdump not available")
    so = os.path.join(ROOT, "gan_b200", "csrc", "libgan_b200.so")
    sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True, timeout=600).stdout
    assert "sm_100a" in sass
    for mnemonic in ("UTCHMMA.2CTA", "UTCHMMA ", "UTMALDG.4D", "UTMALDG.3D", "UTMALDG.2D.2CTA", "UTCBAR.2CTA.MULTICAST", "LDTM."):
        assert mnemonic in sass, mnemonic
    import re
    assert not re.search(r"\bHMMA\.", sass), "mma.sync tensor-core instructions found: the convolutions must be tcgen05"
