#!/bin/bash
# compute-sanitizer over the fast-path tests that exercise the round-end changes (tag = $1): split level loops at every
# geometry (nlay 40 / 47 / 65 / 72 straddle the albedo split), the small-table pointer table with nothing / part / all of
# the block staged, the draw-free McICA path, two slot groups, a few randomized configurations
T='fast_path_runtime_nlay_f32 or partial_table_staging or binary_cloud or two_minor_groups or randomized_configurations[1] or randomized_configurations[7] or randomized_configurations[11] or randomized_configurations[19]'
O=gpurun_out/$1_sanitizer.txt
echo "## memcheck" > $O
compute-sanitizer --tool memcheck --error-exitcode 1 python -m pytest tests -m gpu -x -q -k "$T" 2>&1 | grep -E "passed|failed|ERROR SUMMARY|Invalid|error" | tail -5 >> $O
echo "## racecheck" >> $O
compute-sanitizer --tool racecheck python -m pytest tests -m gpu -x -q -k "fast_path_runtime_nlay_f32 or binary_cloud or partial_table_staging[6000]" 2>&1 | grep -E "passed|failed|RACECHECK SUMMARY|hazard" | tail -5 >> $O
cat $O
