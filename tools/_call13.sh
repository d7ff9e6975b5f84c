export PYTHONUNBUFFERED=1
python tests/golden/make_schedule_diffs.py gpurun_out/schedule_dependent_diffs.json 2>&1 | tail -12
python tests/golden/make_schedule_diffs.py gpurun_out/schedule_dependent_diffs_2.json > /dev/null 2>&1
cmp gpurun_out/schedule_dependent_diffs.json gpurun_out/schedule_dependent_diffs_2.json && echo "second mint identical"
