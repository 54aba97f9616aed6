bash tools/gpu_call.sh r2at smoke tests anchor sanitize bench ref ncu_shade
