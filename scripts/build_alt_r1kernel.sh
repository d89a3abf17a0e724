#!/bin/bash
# Same-box A/B reference: the CURRENT tree (ABI, operators, small kernels) with the ROUND-1 Engine F kernel
# (git show b958506:ursonet_b200/csrc/conv_gemm.cu) and the halo policy off -> ursonet_b200/alt_r1kernel.so
set -e
root=$(cd "$(dirname "$0")/.." && pwd)
tmp=$(mktemp -d)
mkdir -p $tmp/ursonet_b200 $tmp/include
cp -r $root/ursonet_b200/csrc $tmp/ursonet_b200/
cp $root/include/urso_b200.h $tmp/include/
git -C $root show b958506:ursonet_b200/csrc/conv_gemm.cu > $tmp/ursonet_b200/csrc/conv_gemm.cu
cat >> $tmp/ursonet_b200/csrc/conv_gemm.cu <<'EOC'
extern "C" int urso_convgemm_plan_info(const urso_convgemm_t* h, int32_t* out9) {
  const int32_t v[9] = {h->block_n, 1, h->params.stages, h->params.kpack, h->params.halo, 0, h->params.a_stages, h->smem_bytes, h->grid};
  for (int i = 0; i < 9; ++i) out9[i] = v[i];
  return 0;
}
EOC
sed -i 's/  return halo \* 100 <= best \* 107;/  (void)halo; (void)best; return false;/' $tmp/ursonet_b200/csrc/conv_ops.cu
rm -f $tmp/ursonet_b200/csrc/*.o
make -C $tmp/ursonet_b200/csrc -j8 > /dev/null 2>&1
cp $tmp/ursonet_b200/liburso_b200.so $root/ursonet_b200/alt_r1kernel.so
rm -rf $tmp
echo "built ursonet_b200/alt_r1kernel.so"
