mkdir -p gpurun_out
rm -f gpurun_out/parity_report.json
timeout 1800 python -m pytest tests -m gpu -q 2>&1 | tail -40 > gpurun_out/r2_pytest5.log
tail -30 gpurun_out/r2_pytest5.log
