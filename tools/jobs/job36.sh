( time bash tools/final_check.sh ) 2>&1 | tee gpurun_out/final_check_lease2.log
python bench.py --impl reference --steps 3 --warmup 1 2>/dev/null | tail -1 | cut -c1-300 | tee -a gpurun_out/final_check_lease2.log
