#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_engine.py -q -x -m gpu 2>&1 | tail -3
{
TAG=tapblk python scratch/time_fwd.py
TAG=tapblk_noMMA DL4DS_HALO_DBG=1 python scratch/time_fwd.py
TAG=tapblk_skeleton DL4DS_HALO_DBG=7 python scratch/time_fwd.py
TAG=tapblk_tma DL4DS_HALO_TMA=1 python scratch/time_fwd.py
TAG=tapblk_nores DL4DS_HALO_NO_RESIDENT=1 python scratch/time_fwd.py
} 2>&1 | grep -v Warning | tee gpurun_out/r02d_time_fwd6.log
python scratch/halo_stamps.py 64 64 64 48 32 3 2 2>&1 | grep -v Warn | head -14 | tee gpurun_out/r02d_halo_stamps3.log
