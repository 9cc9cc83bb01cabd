#!/usr/bin/env python
"""run.py -- scene orchestration with the reference's flags (run.py:11-44) around the B200 `apd` binary.

    python run.py --data_dir <root> --scans scan1 scan2 --gpu_num 8 --work_num 1 [--only_fuse] [--no_fuse] ...

One `apd` process per scan, pinned to a GPU slot exactly like the reference's worker table (run.py:72-82, 104-138).
Differences: running SAM is out of scope (tools/run_SAM.py is never imported): label maps already present in
<scan>/sa_masks/ are consumed (--use_sa true) unless --no_sam is given, a scan without them runs with --use_sa false
instead of calling the SAM runner (run.py:94-98); --APD_path defaults to the in-tree build.  The image-layout helper (scripts/dataset_loader.py:54-140: first existing candidate
directory, symlinked to <scan>/images) is restated in resolve_images_dir().
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
    p.add_argument('--gpus_per_scan', type=int, default=1,
                   help='extension: one scan over this many GPUs (apd --gpus N: views dealt out in blocks, depth maps exchanged over '
                        'NCCL); --gpu_num GPUs then make gpu_num // gpus_per_scan slots')
    p.add_argument('--scans', type=str, nargs='+', default=[])
    p.add_argument('--only_fuse', action='store_true', default=False)
    p.add_argument('--no_fuse', action='store_true', default=False)
    p.add_argument('--memory_cache', action='store_true', default=False)
    p.add_argument('--no_sam', action='store_true', default=False)
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
    p.add_argument('--reservation', type=str, default=None, help='sleep before starting, e.g. 3h30m10s (run.py:167-170)')
    p.add_argument('--backup_code', action='store_true', default=False, help='copy the sources next to the results (run.py:139)')
    p.add_argument('--image_dir_name', type=str, nargs='+', default=['images', 'undist/images'],
                   help='candidate image directories below the scan folder, first hit wins')
    p.add_argument('--image_suffixes', type=str, nargs='+', default=['.jpg', '.jpeg', '.png'],
                   help='image file extensions that count (with or without the dot, any case)')
    p.add_argument('--no_image_symlink', action='store_true', default=False,
                   help='do not symlink the resolved candidate to <scan>/images')
    return p.parse_args(argv)


def resolve_images_dir(scan_dir, candidates, create_symlink=True):
    """scripts/dataset_loader.py:75-140: the first existing candidate is the source; <scan>/images is what the binary reads, so
    a different source is symlinked there unless told otherwise.  Returns the directory to count images in."""
    source = None
    for cand in candidates:
        path = os.path.join(scan_dir, *[x for x in cand.replace('\\', '/').split('/') if x])
        if os.path.isdir(path):
            source = path
            break
    if source is None:
        raise FileNotFoundError('no image directory among %s below %s' % (candidates, scan_dir))
    canonical = os.path.join(scan_dir, 'images')
    if os.path.isdir(canonical):
        return canonical
    if os.path.exists(canonical):
        raise FileExistsError('%s exists but is not a directory' % canonical)
    if not create_symlink:
        return source
    os.symlink(os.path.relpath(source, scan_dir), canonical)
    return canonical


def count_images(image_dir, suffixes):
    sfx = tuple(('.' + x.lstrip('.')).lower() for x in suffixes)
    return sum(1 for f in os.listdir(image_dir) if f.lower().endswith(sfx))


def parse_reservation(text):
    """'3h30m10s' -> seconds (the reference hands the string to sleep(1))"""
    import re
    total, pos = 0.0, 0
    for m in re.finditer(r'(\d+(?:\.\d+)?)([smhd]?)', text):
        if m.start() != pos or not m.group(0):
            break
        total += float(m.group(1)) * {'': 1, 's': 1, 'm': 60, 'h': 3600, 'd': 86400}[m.group(2)]
        pos = m.end()
    if pos != len(text):
        raise ValueError('bad reservation %r' % text)
    return total


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
    use_sa = not args.no_sam and os.path.isdir(os.path.join(scan_dir, 'sa_masks'))  # run.py:94-98 without the SAM runner
    cmd = [args.APD_path, '--dense_folder', scan_dir, '--gpu_index', str(gpu_index), '--dataset', dataset_tag(args.data_dir, scan),
           '--only_fuse', b(args.only_fuse), '--no_fuse', b(args.no_fuse), '--use_sa', b(use_sa), '--memory_cache', b(args.memory_cache),
           '--flush', b(args.flush), '--export_anchor', b(args.export_anchor), '--export_curve', b(args.export_curve),
           '--export_color', b(not args.no_color), '--use_impetus', b(not args.no_impetus), '--weak_filter', b(not args.no_weak_filter)]
    if args.gpus_per_scan > 1:  # the slot's GPUs: gpu_index .. gpu_index + gpus_per_scan - 1
        cmd += ['--gpus', str(args.gpus_per_scan), '--gpu_list', ','.join(str(gpu_index + g) for g in range(args.gpus_per_scan))]
    return cmd


def _init(pp, ll, aa):
    global positions, lock, args
    positions, lock, args = pp, ll, aa


def worker(scan):
    scan_dir = os.path.join(args.data_dir, scan)
    try:
        resolve_images_dir(scan_dir, args.image_dir_name, not args.no_image_symlink)
    except (FileNotFoundError, FileExistsError) as exc:
        print('[{}] cannot prepare the image directory: {}'.format(scan, exc))
        return 1
    if not os.path.isdir(os.path.join(scan_dir, 'images')):
        print('{} has no images/ folder (and --no_image_symlink was given)'.format(scan_dir))
        return 1
    with lock:
        pos_index = 0
        for j in range(len(positions)):
            if positions[j] == 0:
                positions[j] = 1
                pos_index = j
                break
    try:
        gpu_index = (pos_index // args.work_num) * args.gpus_per_scan
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
    if a.reservation is not None:
        import time
        print('sleep for reservation: {}'.format(a.reservation))
        time.sleep(parse_reservation(a.reservation))
    scans = a.scans or sorted(d for d in os.listdir(a.data_dir) if os.path.isdir(os.path.join(a.data_dir, d)))
    if a.backup_code and not (a.review or a.dry_run):
        import shutil
        import time
        dst = os.path.join(a.data_dir, 'code_backup_' + time.strftime('%Y%m%d_%H%M%S'))
        os.makedirs(dst, exist_ok=True)
        for item in ('run.py', 'include', 'apde_mvs_b200'):
            src = os.path.join(ROOT, item)
            if os.path.isdir(src):
                shutil.copytree(src, os.path.join(dst, item), ignore=shutil.ignore_patterns('_build', '__pycache__', '*.so', '*.o'))
            else:
                shutil.copy(src, dst)
    # scans whose images cannot be found are skipped; most images first, like run.py:186-214
    counted = []
    for s in scans:
        scan_dir = os.path.join(a.data_dir, s)
        if not os.path.isdir(scan_dir):
            print('{} is not a dir'.format(scan_dir))
            continue
        try:
            counted.append((count_images(resolve_images_dir(scan_dir, a.image_dir_name, not a.no_image_symlink), a.image_suffixes), s))
        except (FileNotFoundError, FileExistsError) as exc:
            print('skipping {}: {}'.format(scan_dir, exc))
    if not counted:
        print('No valid scans found.')
        return 0
    counted.sort(key=lambda x: -x[0])
    scans = [s for _, s in counted]
    if a.gpus_per_scan < 1 or a.gpu_num % a.gpus_per_scan:
        print('--gpu_num must be a multiple of --gpus_per_scan')
        return 1
    total = (a.gpu_num // a.gpus_per_scan) * a.work_num
    positions = mp.Array('i', [0] * total)
    lock = mp.Lock()
    with mp.Pool(total, initializer=_init, initargs=(positions, lock, a)) as pool:
        rcs = pool.map(worker, scans)
    return 0 if all(rc == 0 for rc in rcs) else 1


if __name__ == '__main__':
    sys.exit(main())
