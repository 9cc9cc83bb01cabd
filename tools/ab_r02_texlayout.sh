set -x
python tools/dump_maps.py /tmp/m_base.npz 2>&1 | tail -1
APDE_TILE_SHIFT=5 python tools/dump_maps.py /tmp/m_t5.npz 2>&1 | tail -1
python tools/compare_maps.py /tmp/m_base.npz /tmp/m_t5.npz
APDE_LIB=$PWD/ab/pitch/libapde.so python tools/dump_maps.py /tmp/m_pitch.npz 2>&1 | tail -1
python tools/compare_maps.py /tmp/m_base.npz /tmp/m_pitch.npz
bash tools/ab_bench.sh "APDE_TILE_SHIFT=3"
bash tools/ab_bench.sh "APDE_TILE_SHIFT=4"
bash tools/ab_bench.sh "APDE_TILE_SHIFT=5"
bash tools/ab_bench.sh "APDE_TILE_SHIFT=6"
bash tools/ab_bench.sh "APDE_LIB=$PWD/ab/pitch/libapde.so APDE_TILE_SHIFT=3"
bash tools/ab_bench.sh "APDE_LIB=$PWD/ab/pitch/libapde.so APDE_TILE_SHIFT=5"
bash tools/ab_bench.sh "APDE_LIB=$PWD/ab/pitch/libapde.so APDE_TILE_SHIFT=6"
