for st in 3 4 5; do CATB200_GEMM_STAGES=$st python tools/mb_timeline.py 2>&1 | tail -12 > gpurun_out/r2l_timeline_st$st.txt; done
tail -12 gpurun_out/r2l_timeline_st3.txt; tail -12 gpurun_out/r2l_timeline_st5.txt
