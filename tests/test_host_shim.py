"""The C++ host side: drop-in FluidSystemSPH / Grid headers (sph-erosion_b200/host/) over the C ABI,
driven by the headless driver that mirrors the call pattern of the reference's main.cpp."""
import os
import subprocess

import numpy as np
import pytest

from conftest import GOLDEN, ROOT

HOST = os.path.join(ROOT, "sph-erosion_b200", "host")
EXE = os.path.join(HOST, "headless")


def build():
    subprocess.check_call(["make", "-s", "-C", HOST, "headless"])
    assert os.path.exists(EXE)


def read_dump(path):
    raw = np.fromfile(path, np.uint8)
    n = int(raw[:4].view(np.int32)[0])
    f = raw[4:].view(np.float32)
    return f[:3 * n].reshape(n, 3), f[3 * n:6 * n].reshape(n, 3), f[6 * n:7 * n]


def test_headless_builds_and_refuses_to_run_without_gpu():
    import torch
    build()
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    r = subprocess.run([EXE, "--steps", "1"], capture_output=True, text=True)
    assert r.returncode != 0
    assert "no CPU fallback" in r.stderr


def test_shim_compiles_against_the_reference_glm():
    glm = "/root/reference/vendor/glm"
    if not os.path.isdir(glm):
        pytest.skip("reference tree absent")
    subprocess.check_call(["g++", "-std=c++17", "-fsyntax-only", "-DSPHE_USE_GLM", "-I" + glm, "-I" + os.path.join(ROOT, "include"),
                           "-I" + HOST, os.path.join(HOST, "headless_main.cpp")])


@pytest.mark.gpu
def test_headless_default_scene_matches_reference_golden(tmp_path):
    build()
    g = np.load(os.path.join(GOLDEN, "default_scene.npz"))
    out = str(tmp_path / "s1.bin")
    r = subprocess.run([EXE, "--steps", "1", "--dump", out], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    pos, vel, rho = read_dump(out)
    for got, want in ((pos, g["s1_pos"]), (vel, g["s1_vel"]), (rho, g["s1_density"])):
        assert np.abs(got - want).max() <= 1e-5 * np.abs(want).max()
    assert "particles 1000" in r.stdout
    out = str(tmp_path / "s20.bin")
    r = subprocess.run([EXE, "--steps", "20", "--add-at", "10", "--dump", out], capture_output=True, text=True)
    assert r.returncode == 0 and "particles 1125" in r.stdout
    pos, vel, rho = read_dump(out)
    assert pos.shape == (1125, 3) and np.isfinite(pos).all() and rho.min() > 100


@pytest.mark.gpu
def test_headless_terrain_and_erosion(tmp_path):
    build()
    gold = np.load(os.path.join(GOLDEN, "terrain.npz"))
    img = np.zeros((512, 512), np.uint8); img[:64, :64] = gold["hf64"]
    raw = str(tmp_path / "hf.bin"); img.tofile(raw)
    r = subprocess.run([EXE, "--steps", "60", "--terrain", raw, "--erosion", "--dump", str(tmp_path / "e.bin")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert "terrain: 15000 surface floats, 14406 indices, H(5,7) = %d" % int(gold["hf64"][5, 7]) in r.stdout
    pos, vel, rho = read_dump(str(tmp_path / "e.bin"))
    assert np.isfinite(pos).all()


@pytest.mark.gpu
def test_headless_checkpoint_resume_is_bit_exact(tmp_path):
    """--save / --load (FluidSystemSPH::Save / Load -> sphe_save_state / sphe_load_state): 30 steps, checkpoint, 30 more
    steps in a second process == 60 steps in one process, bit for bit, terrain erosion included."""
    build()
    gold = np.load(os.path.join(GOLDEN, "terrain.npz"))
    img = np.zeros((512, 512), np.uint8); img[:64, :64] = gold["hf64"]
    raw = str(tmp_path / "hf.bin"); img.tofile(raw)
    common = ["--terrain", raw, "--erosion"]
    full, part, ck = str(tmp_path / "full.bin"), str(tmp_path / "part.bin"), str(tmp_path / "ck.sphe")
    r = subprocess.run([EXE, "--steps", "60", "--dump", full] + common, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([EXE, "--steps", "30", "--save", ck] + common, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([EXE, "--steps", "30", "--load", ck, "--dump", part] + common, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    a, b = np.fromfile(full, np.uint8), np.fromfile(part, np.uint8)
    assert a.size == b.size and np.array_equal(a, b), "resumed run differs from the uninterrupted one"


SLABS_EXE = os.path.join(HOST, "headless_slabs")


def test_headless_slabs_builds_and_refuses_to_run_without_gpu():
    import torch
    subprocess.check_call(["make", "-s", "-C", HOST, "headless_slabs"])
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    r = subprocess.run([SLABS_EXE, "--slabs", "2"], capture_output=True, text=True)
    assert r.returncode != 0 and "no CPU fallback" in r.stderr


@pytest.mark.gpu
@pytest.mark.parametrize("K", [2, 4])
def test_headless_slabs_cpp_driver_matches_single_handle(K):
    """host/headless_slabs.cpp: K x-slabs driven from ONE C++ thread through the C ABI (peer mailboxes, ring closure for
    K >= 3), census complete and positions BIT-EQUAL to the single-handle run of the same scene."""
    subprocess.check_call(["make", "-s", "-C", HOST, "headless_slabs"])
    r = subprocess.run([SLABS_EXE, "--slabs", str(K), "--axis", "24", "--steps", "40", "--check"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "HEADLESS_SLABS OK" in r.stdout and "bit-equal to the single-handle run: yes" in r.stdout
