#!/bin/bash
# Build liburso_b200.so of another git revision for same-box A/B runs:  scripts/build_ref.sh <git-ref> <name>
# -> ursonet_b200/alt_<name>.so (git-ignored, travels with gpurun); select with URSO_LIB_PATH=ursonet_b200/alt_<name>.so
set -e
ref=$1; name=$2
root=$(cd "$(dirname "$0")/.." && pwd)
tmp=$(mktemp -d)
git -C "$root" archive "$ref" ursonet_b200/csrc include | tar -x -C "$tmp"
make -C "$tmp/ursonet_b200/csrc" -j8 > /dev/null 2>&1
cp "$tmp/ursonet_b200/liburso_b200.so" "$root/ursonet_b200/alt_$name.so"
rm -rf "$tmp"
echo "built ursonet_b200/alt_$name.so from $ref"
