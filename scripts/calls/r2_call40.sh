BE="python bench.py --only c3 --steps 1 --warmup 3 --no-cpu-baseline --no-peaks --no-est"
SKNNR_B200_TRACE=1 timeout 600 $BE > gpurun_out/tr_1m.log 2> gpurun_out/tr_1m.err; grep -A40 "sknnr trace" gpurun_out/tr_1m.err | tail -18
SKNNR_B200_TRACE=1 timeout 600 $BE --chunk-rows 524288 > gpurun_out/tr_512.log 2> gpurun_out/tr_512.err; grep -A40 "sknnr trace" gpurun_out/tr_512.err | tail -26
