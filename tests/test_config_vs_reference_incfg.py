"""The drop-in executable's configuration surface against the REFERENCE's own parser: oracle/_ref/incfg_ref is
ext/incfg/incfg.cpp of the reference, compiled in place by oracle/build_ref.sh, with the option set of the reference's
wass_stereo ("every key in Appendix B, incfg syntax and error behaviour").  Skipped when the binary has
not been built (no reference tree and no prebuilt copy)."""
import os
import subprocess
import pytest
from helpers import ROOT

REF = os.path.join(ROOT, "oracle", "_ref", "incfg_ref")
EXE = os.path.join(ROOT, "wass_b200", "bin", "wass_stereo")
EXT_BLOCKS = ("# B200 extension: use the 8-path (MODE_HH) aggregation instead of the reference's 5-path MODE_SGBM\n"
              "# \n#SGM_FULL_8PATH=false\n\n",
              "# B200 extension: write the diagnostic JPEGs (stereo.jpg, disparity_*.jpg, graph_components.jpg) as the reference "
              "always does\n# \n#SAVE_DEBUG_IMAGES=true\n\n")


@pytest.fixture(scope="module", autouse=True)
def _built():
    if os.path.isdir("/root/reference"):
        subprocess.run(["bash", os.path.join(ROOT, "oracle", "build_ref.sh")], check=True, capture_output=True)
    if not os.path.exists(REF):
        pytest.skip("oracle/_ref/incfg_ref not built")
    from wass_b200 import build
    build.build()


def _ours_without_extension(text):
    for b in EXT_BLOCKS:
        assert b in text
        text = text.replace(b, "")
    return text


def test_genconfig_is_the_reference_string(tmp_path):
    ref = subprocess.run([REF, "--genconfig"], capture_output=True, text=True)
    assert ref.returncode == 0
    r = subprocess.run([EXE, "--genconfig"], cwd=tmp_path, capture_output=True, text=True)
    assert r.returncode == 0
    assert _ours_without_extension((tmp_path / "stereo_config.txt").read_text()) == ref.stdout


SCENARIOS = [
    "",                                                                     # empty file: all defaults
    "# only a comment\n\n\n",
    "WINSIZE=11\nMAX_DISPARITY = 256\n",
    "  WINSIZE = 11 \r\nLEFT_MASK_IMAGE = \"my mask.png\"\r\n\r\nSAVE_AS_PLY=true\n",
    "DENSE_SCALE=0.5\nPLANE_MAX_DISTANCE=1.25e0\nZGAP_PERCENTILE=98.5\n",
    "SAVE_AS_PLY=false\nDISABLE_AUTO_LEFT_RIGHT=true\nSWAP_LEFT_RIGHT=true\n",
    "RANDOM_SEED=-1\nDISPARITY_OFFSET=-3\nMIN_DISPARITY=1\n",
    "WINSIZE=11 # trailing comment\n",
    "WINSIZE=11\nWINSIZE=13\n",                                              # duplicate key
    "LEFT_MASK_IMAGE=mask.png\n",                                           # unquoted string
    "LEFT_MASK_IMAGE=\"a # b.png\"\n",                                       # '#' inside quotes
    "LEFT_MASK_IMAGE=\"\"\n",
    "SAVE_AS_PLY=yes\n",                                                    # bad boolean
    "SAVE_AS_PLY=TRUE\n",
    "SAVE_AS_PLY=1\n",
    "WINSIZE=abc\n",                                                        # bad integer
    "WINSIZE=11.5\n",
    "WINSIZE=\n",
    "WINSIZE\n",                                                            # no '='
    "=11\n",
    "NOT_A_KEY=3\n",                                                        # unknown key
    "winsize=11\n",                                                         # keys are case sensitive
    "MAX_DISPARITY=256=7\n",
    "PLANE_MAX_DISTANCE=abc\n",
    "PLANE_MAX_DISTANCE=1,5\n",
    "\tWINSIZE\t=\t9\t\n",
    "WINSIZE=9",                                                            # no final newline
    "TRIANG_BBOX_TOP=-1\nTRIANG_BBOX_LEFT=10\n",
    "USE_CUSTOM_STEREORECTIFY=true\nRECTIFY_ANGLE=2.5\nDISABLE_RECTIFY_ROI=false\n",
]


@pytest.mark.parametrize("idx", range(len(SCENARIOS)))
def test_load_agrees_with_reference_parser(tmp_path, idx):
    text = SCENARIOS[idx]
    cfg = tmp_path / "cfg.txt"
    cfg.write_bytes(text.encode())
    ref = subprocess.run([REF, str(cfg)], capture_output=True, text=True)
    wd = tmp_path / "wd"
    wd.mkdir()
    r = subprocess.run([EXE, str(cfg), str(wd)], capture_output=True, text=True)
    saved = wd / "stereo_config.txt"
    if ref.returncode != 0:
        # the reference aborts on the load error (wass_stereo.cpp:1840-1846): so must the drop-in, before saving anything,
        # and with the parser's own message
        assert ref.stdout.startswith("ERROR: ")
        assert r.returncode == 255 and not saved.exists(), (text, r.stdout[-500:])
        assert ref.stdout[len("ERROR: "):].strip() in r.stdout, (ref.stdout, r.stdout[-800:])
    else:
        # parsed: the drop-in goes on (and fails later for want of input data), leaving the parsed options in the workdir
        # exactly as the reference would write them (wass_stereo.cpp:1848-1856)
        assert saved.exists(), (text, r.stdout[-500:])
        assert _ours_without_extension(saved.read_text()) == ref.stdout, text
