#!/usr/bin/env bash
# config 5 on one GPU: the device-resident frame pipeline, with and without the exact PICP early-out
cd "$(dirname "$0")/.."
for ne in 0 1; do
  VO_PICP_NO_EARLY_OUT=$ne ./visual-odometry_b200/host/bin/vo_sequence ${1:-100000} ${2:-1000} 1000 100 | cut -c1-330
done
