timeout 600 python scratch/occ_check.py > gpurun_out/occ_check3.log 2>&1; tail -11 gpurun_out/occ_check3.log
