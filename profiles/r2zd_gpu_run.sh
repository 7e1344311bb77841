mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "ghost or gf or peak or c3" 2>&1 | tail -3
python profiles/r2_c3_run.py 2>&1 | tail -1 | cut -c1-200
