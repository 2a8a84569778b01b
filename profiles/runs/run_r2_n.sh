set -x
O=gpurun_out
timeout -s KILL 900 ncu --set full --clock-control none --import-source on -k regex:"k_assemble_rows|k_assemble_prep" -c 2 -o $O/r2n_asm -f python profiles/prof_gs.py > $O/r2n.log 2>&1
python profiles/ncu_summary.py $O/r2n_asm.ncu-rep $O/r2n_ncu_asm.md
python profiles/ncu_source.py $O/r2n_asm.ncu-rep 1 40 > $O/r2n_ncu_source_asm.txt 2>&1
ncu -i $O/r2n_asm.ncu-rep --page details > $O/r2n_details.txt 2>&1
rm -f $O/r2n_asm.ncu-rep
