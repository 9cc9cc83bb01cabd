# identity of the current build against the previous commit's build (ab/prev/libapde.so) over a whole schedule with weak
# texture, then stage times on the headline and on the weak-texture (C3) workload
python tools/dump_maps.py /tmp/m_now.npz 640 480 5 4 0.3 2 2>&1 | tail -1
APDE_LIB=$PWD/ab/prev/libapde.so python tools/dump_maps.py /tmp/m_prev.npz 640 480 5 4 0.3 2 2>&1 | tail -1
python tools/compare_maps.py /tmp/m_now.npz /tmp/m_prev.npz
bash tools/ab_bench.sh "APDE_LIB=$PWD/ab/prev/libapde.so"
bash tools/ab_bench.sh "X=1"
bash tools/ab_bench.sh "APDE_LIB=$PWD/ab/prev/libapde.so" --config C3
bash tools/ab_bench.sh "X=1" --config C3
