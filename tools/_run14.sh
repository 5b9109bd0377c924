mkdir -p gpurun_out/ev
( time python tools/ber_sweep.py ) > gpurun_out/ev/ber_fer.md 2> gpurun_out/ev/ber_fer.err
tail -3 gpurun_out/ev/ber_fer.err; grep -c "| yes |" gpurun_out/ev/ber_fer.md; grep -c "| NO\|| no" gpurun_out/ev/ber_fer.md
