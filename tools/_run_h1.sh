mkdir -p gpurun_out
for tool in memcheck racecheck synccheck; do
  timeout 400 compute-sanitizer --tool $tool --error-exitcode 9 python tools/sanitize.py > gpurun_out/san_$tool.log 2>&1; echo "$tool rc=$?" >> gpurun_out/san_$tool.log; grep -E "rc=|ERROR SUMMARY|iters equal|track ok|Race|hazard" gpurun_out/san_$tool.log | head -12
done
