bash tools/gpu_t3.sh g
python tools/prof_classes.py fp32 2>&1 | tail -8
EDTTS_LIB=$PWD/edge_diffusion_tts_b200/lib/libedtts_clk.so python tools/prof_classes.py fp32 256 400 1 2>&1 | grep t3clk | tail -9
