"""CPU tests of run.py, the scene orchestration with the reference's flags (run.py:11-221): command line handed to the
binary, dataset tags, image-directory resolution, scan ordering."""
import os
import subprocess
import sys

from helpers import ROOT

RUN = os.path.join(ROOT, "run.py")


def _mk(root, scan, sub, names):
    d = os.path.join(root, scan, *sub.split("/"))
    os.makedirs(d)
    for n in names:
        open(os.path.join(d, n), "w").close()


def test_dry_run_commands_layout_and_order(tmp_path):
    root = str(tmp_path / "ETH3D_like")
    _mk(root, "small", "undist/images", ["00000000.JPG"])                       # alternative layout -> symlinked to images/
    _mk(root, "big", "images", ["%08d.png" % i for i in range(4)] + ["notes.txt"])
    os.makedirs(os.path.join(root, "empty"))
    out = subprocess.run([sys.executable, RUN, "--data_dir", root, "--dry_run", "--gpu_num", "1", "--no_weak_filter", "--no_color"],
                         capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    lines = [l for l in out.stdout.splitlines() if "--dense_folder" in l]
    assert len(lines) == 2 and "skipping" in out.stdout            # the scan without images is skipped (run.py:186-205)
    assert "big" in lines[0] and "small" in lines[1]                # most images first (run.py:214); the .txt does not count
    assert os.path.islink(os.path.join(root, "small", "images"))
    assert os.listdir(os.path.join(root, "small", "images")) == ["00000000.JPG"]
    for l in lines:  # booleans take a value, spelled as main.cpp:9-24 expects them
        assert "--dataset ETH3D" in l and "--weak_filter false" in l and "--export_color false" in l and "--use_impetus true" in l
        assert "--use_sa false" in l and "--only_fuse false" in l
    # label maps of tools/run_SAM.py already in the scan are consumed unless --no_sam (run.py:94-98, 113)
    os.makedirs(os.path.join(root, "big", "sa_masks"))
    for extra, want in (([], "--use_sa true"), (["--no_sam"], "--use_sa false")):
        out = subprocess.run([sys.executable, RUN, "--data_dir", root, "--scans", "big", "--dry_run"] + extra, capture_output=True, text=True)
        assert want in out.stdout, out.stdout
    # --no_image_symlink: the alternative layout is left alone and the scan cannot run
    root2 = str(tmp_path / "other")
    _mk(root2, "s", "undist/images", ["00000000.jpg"])
    out = subprocess.run([sys.executable, RUN, "--data_dir", root2, "--dry_run", "--no_image_symlink"], capture_output=True, text=True)
    assert not os.path.exists(os.path.join(root2, "s", "images")) and "--dense_folder" not in out.stdout


def test_gpus_per_scan_slots(tmp_path):
    """--gpus_per_scan N: gpu_num // N slots, every command names its slot's devices for `apd --gpus N`"""
    root = str(tmp_path / "scenes")
    for k in range(3):
        _mk(root, "scan%d" % k, "images", ["%08d.png" % i for i in range(3 - k)])
    out = subprocess.run([sys.executable, RUN, "--data_dir", root, "--dry_run", "--gpu_num", "8", "--gpus_per_scan", "4"], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    lines = [l for l in out.stdout.splitlines() if "--dense_folder" in l]
    assert len(lines) == 3
    lists = sorted(l.split("--gpu_list ")[1].split()[0] for l in lines)
    assert set(lists) <= {"0,1,2,3", "4,5,6,7"} and all("--gpus 4" in l for l in lines)
    for l in lines:
        assert "--gpu_index %s " % l.split("--gpu_list ")[1].split(",")[0] in l
    bad = subprocess.run([sys.executable, RUN, "--data_dir", root, "--dry_run", "--gpu_num", "6", "--gpus_per_scan", "4"], capture_output=True, text=True)
    assert bad.returncode != 0 and "multiple of" in bad.stdout
    one = subprocess.run([sys.executable, RUN, "--data_dir", root, "--dry_run", "--gpu_num", "2"], capture_output=True, text=True)
    assert "--gpus" not in one.stdout.replace("--gpus_per_scan", "")


def test_dataset_tags_and_reservation():
    sys.path.insert(0, ROOT)
    import importlib
    run = importlib.import_module("run")
    assert run.dataset_tag("/data/DTU/test", "scan1") == "DTU"
    assert run.dataset_tag("/data/TaT", "Ballroom") == "TaT_a" and run.dataset_tag("/data/TaT", "Family") == "TaT_i"
    assert run.dataset_tag("/data/ETH3D", "office") == "ETH3D" and run.dataset_tag("/data/mine", "x") == "General"
    assert run.parse_reservation("3h30m10s") == 3 * 3600 + 30 * 60 + 10 and run.parse_reservation("15") == 15
