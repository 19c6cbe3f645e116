set -x
python scratch/prof_sparse.py || exit 1
ncu --set full --clock-control none --import-source on -k regex:tapgrad_sparse -s 2 -c 1 -o gpurun_out/prof_sparse_r1 -f python scratch/prof_sparse.py > gpurun_out/ncu_sparse.log 2>&1
DENS=0.0 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"tapgrad_sparse|reduce_partials" --csv --log-file gpurun_out/sparse_d0.csv python scratch/prof_sparse.py > /dev/null 2>&1
DENS=0.016 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"tapgrad_sparse|reduce_partials" --csv --log-file gpurun_out/sparse_d16.csv python scratch/prof_sparse.py > /dev/null 2>&1
tail -3 gpurun_out/ncu_sparse.log
