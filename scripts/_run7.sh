mkdir -p gpurun_out/r2g
rm -f gpurun_out/r2g/abl.jsonl
for a in 16 0 256 0 256; do NB_TC_ABLATE=$a timeout 300 python scripts/abl_probe.py >> gpurun_out/r2g/abl.jsonl 2>> gpurun_out/r2g/abl.err; done
