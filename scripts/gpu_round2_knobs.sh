#!/bin/bash
# First GPU call of the next round: A/B of the knobs that are implemented but not yet measured.
mkdir -p gpurun_out
{
for push in 0 6 12 50; do B200AG_ROLL_PUSH=$push timeout 200 python scripts/replay_config.py c4 20 5 2>&1 | tail -1; done
for reuse in 2 4; do B200AG_REUSE_LIMIT=$reuse B200AG_ROLL_PUSH=12 timeout 200 python scripts/replay_config.py c4 20 5 2>&1 | tail -1; done
for chunk in 0 1; do B200AG_ROW_CHUNKED=$chunk B200AG_ROLL_PUSH=12 timeout 200 python scripts/replay_config.py c4 20 5 2>&1 | tail -1; done
for tile in 0 1 3; do for sp in 0 1 2 4; do B200AG_GEMM_TILE=$tile B200AG_GEMM_SPLITS=$sp timeout 200 python scripts/replay_config.py c3 8 5 2>&1 | tail -1; done; done
} | tee gpurun_out/round2_knobs.log
