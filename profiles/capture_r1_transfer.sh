set -x
O=gpurun_out
mkdir -p $O
(timeout -s KILL 600 python -m pytest tests -m gpu -q) 2>&1 | tail -1 > $O/pytest_gpu_tail.log
timeout -s KILL 600 python bench.py > $O/r1_bench_c2.json 2> $O/bench.err
timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:"k_plane2_scatter|k_g2p|k_hessian_gather" -c 8 \
    -o $O/r1_full_transfer -f python bench.py --steps 1 --warmup 3 --cpu-reps 0 > $O/b_full1.log 2>&1
python profiles/ncu_summary.py $O/r1_full_transfer.ncu-rep $O/r1_ncu_full_transfer.md > /dev/null
python profiles/ncu_traffic.py $O/r1_full_transfer.ncu-rep $O/r1_traffic.json > /dev/null
python profiles/ncu_source.py $O/r1_full_transfer.ncu-rep 0 30 > $O/r1_ncu_source_k_plane2_scatter_P2GPolicy_.txt
rm -f $O/r1_full_transfer.ncu-rep
cat $O/pytest_gpu_tail.log
