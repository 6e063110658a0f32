# compute-sanitizer memcheck over every entry point on small ragged cases, incl. the SOLAR_RADVAL build (tools/memcheck_small.py)
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 7 python tools/memcheck_small.py > gpurun_out/t1e_memcheck.log 2>&1; echo "memcheck rc=$?" >> gpurun_out/t1e_memcheck.log
tail -8 gpurun_out/t1e_memcheck.log
