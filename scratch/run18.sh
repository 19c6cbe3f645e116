for k in 9 15; do
python bench.py --workload config4 --kernel $k --steps 10 --warmup 3 > gpurun_out/c4_$k.json 2> gpurun_out/c4_$k.err; echo "rc=$?"; wc -c gpurun_out/c4_$k.json
done
python bench.py --workload config5 --steps 20 --warmup 3 > gpurun_out/c5.json 2> gpurun_out/c5.err; echo "rc=$?"; wc -c gpurun_out/c5.json
grep -v Warning gpurun_out/c5.err | tail -5 | cut -c1-300
