mkdir -p gpurun_out
python tools/debug_env.py Test_03 517 3 > gpurun_out/dbg_default.txt 2>&1
FL_OBS_CTAS=1 python tools/debug_env.py Test_03 517 3 > gpurun_out/dbg_ctas1.txt 2>&1
FL_OBS_TABLES=0 python tools/debug_env.py Test_03 517 3 > gpurun_out/dbg_notables.txt 2>&1
head -50 gpurun_out/dbg_default.txt; echo ======; head -30 gpurun_out/dbg_ctas1.txt; echo =====; head -30 gpurun_out/dbg_notables.txt
