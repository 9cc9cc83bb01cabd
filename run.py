#!/usr/bin/env python
"""run.py -- scene orchestration with the reference's flags (run.py:11-44) around the B200 `apd` binary.

    python run.py --data_dir <root> --scans scan1 scan2 --gpu_num 8 --work_num 1 [--only_fuse] [--no_fuse] ...

One `apd` process per scan, pinned to a GPU slot exactly like the reference's worker table (run.py:72-82, 104-138).
Differences: --no_sam is implied (the SAM plug-in is out of scope, tools/run_SAM.py is never imported), the image layout
helpers of scripts/dataset_loader.py are reduced to "images/ must exist", and --APD_path defaults to the in-tree build.
"""
import argparse
import multiprocessing as mp
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.abspath(__file__))


def parse_args(argv=None):
    p = argparse.ArgumentParser()
    p.add_argument('--data_dir', type=str, required=True)
    p.add_argument('--APD_path', type=str, default=os.path.join(ROOT, 'apde_mvs_b200', '_build', 'apd'))
    p.add_argument('--resume', action='store_true', default=False)
    p.add_argument('--gpu_num', type=int, default=1)
    p.add_argument('--work_num', type=int, default=1)
    p.add_argument('--scans', type=str, nargs='+', default=[])
    p.add_argument('--only_fuse', action='store_true', default=False)
    p.add_argument('--no_fuse', action='store_true', default=False)
    p.add_argument('--memory_cache', action='store_true', default=False)
    p.add_argument('--no_sam', action='store_true', default=True)
    p.add_argument('--no_impetus', action='store_true', default=False)
    p.add_argument('--no_weak_filter', action='store_true', default=False)
    p.add_argument('--no_color', action='store_true', default=False)
    p.add_argument('--flush', action='store_true', default=False)
    p.add_argument('--dry_run', action='store_true', default=False)
    p.add_argument('--ETH3D_train', action='store_true', default=False)
    p.add_argument('--ETH3D_test', action='store_true', default=False)
    p.add_argument('--TaT_intermediate', action='store_true', default=False)
    p.add_argument('--TaT_advanced', action='store_true', default=False)
    p.add_argument('--export_anchor', action='store_true', default=False)
    p.add_argument('--export_curve', action='store_true', default=False)
    p.add_argument('--review', action='store_true', default=False)
    return p.parse_args(argv)


def dataset_tag(data_dir, scan):
    """run.py:83-92"""
    if data_dir.find('DTU') != -1:
        return 'DTU'
    if data_dir.find('TaT') != -1:
        return 'TaT_a' if scan in ['Auditorium', 'Ballroom', 'Courtroom', 'Museum', 'Palace', 'Temple'] else 'TaT_i'
    if data_dir.find('ETH3D') != -1:
        return 'ETH3D'
    return 'General'


def build_command(args, scan_dir, scan, gpu_index):
    b = lambda v: 'true' if v else 'false'  # noqa: E731
    return [args.APD_path, '--dense_folder', scan_dir, '--gpu_index', str(gpu_index), '--dataset', dataset_tag(args.data_dir, scan),
            '--only_fuse', b(args.only_fuse), '--no_fuse', b(args.no_fuse), '--use_sa', 'false', '--memory_cache', b(args.memory_cache),
            '--flush', b(args.flush), '--export_anchor', b(args.export_anchor), '--export_curve', b(args.export_curve),
            '--export_color', b(not args.no_color), '--use_impetus', b(not args.no_impetus), '--weak_filter', b(not args.no_weak_filter)]


def _init(pp, ll, aa):
    global positions, lock, args
    positions, lock, args = pp, ll, aa


def worker(scan):
    scan_dir = os.path.join(args.data_dir, scan)
    if not os.path.isdir(os.path.join(scan_dir, 'images')):
        print('{} has no images/ folder'.format(scan_dir))
        return 1
    with lock:
        pos_index = 0
        for j in range(len(positions)):
            if positions[j] == 0:
                positions[j] = 1
                pos_index = j
                break
    try:
        gpu_index = pos_index // args.work_num
        apd_dir = os.path.join(scan_dir, 'APD')
        os.makedirs(apd_dir, exist_ok=True)
        if args.resume and os.path.exists(os.path.join(apd_dir, 'APD.ply')):
            print('APD result exists for {}'.format(scan_dir))
            return 0
        cmd = build_command(args, scan_dir, scan, gpu_index)
        print(' '.join(cmd))
        if args.review or args.dry_run:
            return 0
        log_path = os.path.join(apd_dir, 'log.txt')
        with open(log_path, 'a' if os.path.exists(log_path) else 'w') as log:
            return subprocess.call(cmd, stdout=log, stderr=subprocess.STDOUT)
    finally:
        with lock:
            positions[pos_index] = 0


def main(argv=None):
    a = parse_args(argv)
    scans = a.scans or sorted(d for d in os.listdir(a.data_dir) if os.path.isdir(os.path.join(a.data_dir, d)))
    # most images first, like run.py:214
    def nimg(s):
        d = os.path.join(a.data_dir, s, 'images')
        return len(os.listdir(d)) if os.path.isdir(d) else 0
    scans.sort(key=nimg, reverse=True)
    total = a.gpu_num * a.work_num
    positions = mp.Array('i', [0] * total)
    lock = mp.Lock()
    with mp.Pool(total, initializer=_init, initargs=(positions, lock, a)) as pool:
        rcs = pool.map(worker, scans)
    return 0 if all(rc == 0 for rc in rcs) else 1


if __name__ == '__main__':
    sys.exit(main())
