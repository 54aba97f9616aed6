bash tools/gpu_call.sh r2ap launches ncu_trace
