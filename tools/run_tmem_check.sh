# developer script: per-phase profile of the resident kernel + per-episode overhead (1-step fits)
CWT_RESIDENT_TMEM=${1:-1} timeout 300 python tools/prof_resident.py --episodes 64 2>&1 | tail -14
CWT_RESIDENT_TMEM=${1:-1} timeout 300 python tools/prof_resident.py --episodes 64 --iters 1 --prof 0 2>&1 | tail -2
CWT_RESIDENT_TMEM=${1:-1} timeout 300 python tools/prof_resident.py --episodes 64 --iters 100 --prof 0 2>&1 | tail -2
