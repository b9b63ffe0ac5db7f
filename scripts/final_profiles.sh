#!/bin/bash
# Round-end measurement pass on one B200: GPU tests, bench line, ncu launch list and one full capture per hot kernel.
# Output lands in gpurun_out/ (scripts/summarize_profiles.py r02 turns it into profiles/).
set -u
R=${1:-r02}
python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/${R}_pytest_gpu.txt
python bench.py --steps 20 --warmup 5 > gpurun_out/${R}_bench_1gpu.json 2> gpurun_out/${R}_bench_1gpu.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${R}_bench_reference.json 2>> gpurun_out/${R}_bench_1gpu.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_${R}.csv \
    python bench.py --steps 2 --warmup 3 --quick > gpurun_out/ncu_launches.log 2>&1
for k in "csweep_fwd:k_csweep:8" "csweep_adj:k_csweep:29" "sigma:k_sigma$:3"; do
    n=${k%%:*}; r=${k#*:}; sk=${r##*:}; r=${r%:*}
    ncu --set full --clock-control none --import-source on -k "regex:$r" --launch-skip $sk -c 1 -o gpurun_out/${R}_$n -f \
        python bench.py --steps 2 --warmup 3 --quick > gpurun_out/ncu_$n.log 2>&1
done
ncu --set full --clock-control none --import-source on -k regex:k_rdm_gram -c 1 -o gpurun_out/${R}_rdm_gram -f python scripts/rdm24.py > gpurun_out/ncu_rdm.log 2>&1
tail -3 gpurun_out/ncu_rdm.log
cat gpurun_out/${R}_pytest_gpu.txt
