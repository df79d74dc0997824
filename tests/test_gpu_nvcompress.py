"""nvcompress, the reference's own command-line tool (src/nvtt/tools/compress.cpp compiled UNCHANGED), built once against the
B200 drop-in library and once against the reference library (tests/build_nvcompress.sh): the same command line on the same
input file must write byte-identical .dds / .ktx files.  Input files are made here with PIL (PNG) - image-file decoding is
the tool's business (the reference's nvimage reader in both binaries), not the product's."""
import os
import subprocess

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))
OURS = os.path.join(HERE, "_build", "nvcompress_b200")
REF = os.path.join(HERE, "_build", "nvcompress_ref")


@pytest.fixture(scope="module")
def files(nvtt, tmp_path_factory):
    if not (os.path.exists(OURS) and os.path.exists(REF)):
        pytest.skip("tests/_build/nvcompress_{b200,ref} not built (needs /root/reference at build time)")
    from PIL import Image
    d = tmp_path_factory.mktemp("nvcompress")
    s = nvtt.synth
    out = {}

    def png(name, bgra, mode="RGBA"):
        rgba = bgra[..., [2, 1, 0, 3]].copy()
        p = str(d / (name + ".png"))
        (Image.fromarray(rgba, "RGBA") if mode == "RGBA" else Image.fromarray(rgba[..., :3].copy(), "RGB")).save(p)
        out[name] = p

    png("color", s.photo_bgra8(256, 256, seed=1234), mode="RGB")
    png("alpha", s.photo_bgra8(256, 128, seed=99, alpha=True))
    png("normal", s.normal_bgra8(128, 128, seed=7), mode="RGB")
    png("odd", s.photo_bgra8(100, 60, seed=3, alpha=True))
    png("large", s.photo_bgra8(2048, 2048, seed=5), mode="RGB")
    out["dir"] = str(d)
    return out


def _run(binary, args, src, dst):
    env = dict(os.environ)
    r = subprocess.run([binary, "-silent"] + args + [src, dst], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, env=env, timeout=600)
    assert r.returncode == 0, (binary, args, r.stdout.decode(errors="replace")[-2000:])
    return open(dst, "rb").read()


CASES = [
    ("color", ["-bc1"]), ("large", ["-bc1"]), ("color", ["-bc1", "-fast"]), ("color", ["-bc1", "-nomips"]), ("color", ["-bc1", "-srgb", "-dds10"]),
    ("alpha", ["-bc3", "-alpha", "-mipfilter", "kaiser"]), ("alpha", ["-bc3", "-mipfilter", "triangle", "-repeat"]),
    ("alpha", ["-bc1a", "-alpha"]), ("alpha", ["-bc2", "-alpha"]), ("alpha", ["-bc3", "-fast", "-clamp"]),
    ("normal", ["-bc5", "-normal"]), ("normal", ["-bc3n", "-normal"]), ("color", ["-bc5", "-tonormal"]),
    ("color", ["-bc4"]), ("color", ["-bc1", "-ktx"]), ("alpha", ["-bc3", "-alpha", "-ktx", "-mipfilter", "kaiser"]),
    ("odd", ["-bc7", "-dds10"]), ("odd", ["-bc1"]), ("odd", ["-bc3", "-alpha", "-mipfilter", "kaiser"]),
    ("color", ["-rgb"]), ("alpha", ["-rgb", "-alpha"]), ("color", ["-lumi"]),
    ("odd", ["-bc6"]),                 # image path: Surface::load -> context.compress(image, ...)
    ("odd", ["-bc3_rgbm"]),            # Surface::load, range, scaleBias, toneMap, clamp, toGamma(2)
    ("odd", ["-bc3", "-rgbm"]),        # ... + toRGBM
    ("odd", ["-bc3_rgbm", "-rangescale"]),
]


@pytest.mark.parametrize("name,args", CASES, ids=lambda v: v if isinstance(v, str) else "_".join(a.strip("-") for a in v))
def test_nvcompress_byte_identical(files, name, args):
    src = files[name]
    ext = ".ktx" if "-ktx" in args else ".dds"
    tag = name + "_" + "_".join(a.strip("-") for a in args)
    want = _run(REF, args, src, os.path.join(files["dir"], tag + "_ref" + ext))
    got = _run(OURS, args, src, os.path.join(files["dir"], tag + "_b200" + ext))
    assert len(got) == len(want), (len(got), len(want))
    if got != want:
        a, b = np.frombuffer(got, np.uint8), np.frombuffer(want, np.uint8)
        bad = np.nonzero(a != b)[0]
        raise AssertionError("%s %s: %d of %d bytes differ, first at %d" % (name, args, bad.size, a.size, int(bad[0])))


def test_nvcompress_dds_input_with_premade_mips(files):
    """A .dds with its own mip chain as INPUT (tools/compress.cpp:506-557: every surface goes to InputOptions::setMipmapData)."""
    src = os.path.join(files["dir"], "premade.dds")
    _run(REF, ["-rgb", "-mipfilter", "kaiser"], files["alpha"], src)  # BGRA8 .dds with a Kaiser chain
    for args in (["-bc1"], ["-bc3", "-alpha"], ["-bc1", "-nomips"]):
        tag = "premade_" + "_".join(a.strip("-") for a in args)
        want = _run(REF, args, src, os.path.join(files["dir"], tag + "_ref.dds"))
        got = _run(OURS, args, src, os.path.join(files["dir"], tag + "_b200.dds"))
        assert got == want, args
