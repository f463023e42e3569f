#!/bin/bash
# usage: tools/ab_run.sh "<bench args>" tag1 tag2 ...   -- bench_brief over A/B builds (zkvm_b200/libzkmsm_<tag>.so; "base" = the product build)
args=$1; shift
for tag in "$@"; do
  lib=zkvm_b200/libzkmsm_$tag.so; [ "$tag" = base ] && lib=zkvm_b200/libzkmsm.so
  bash tools/bench_brief.sh $PWD/$lib $args
done
