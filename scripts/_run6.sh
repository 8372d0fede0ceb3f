mkdir -p gpurun_out/r2f
NB_TC_PROF=1 timeout 300 python scripts/abl_probe.py > gpurun_out/r2f/prof.jsonl 2> gpurun_out/r2f/prof.err
grep "nb_tc prof" gpurun_out/r2f/prof.err | sort | uniq -c | sort -rn | head -20 > gpurun_out/r2f/prof_summary.txt
