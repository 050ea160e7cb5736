#!/bin/bash
# round 2, session V: packed-RGB kernels: resident CTAs per SM and rows per CTA; ncu --set full of the five flavours
mkdir -p gpurun_out
for rep in 1 2; do
python tools/rgb24_ab.py --label "product (6/6/6 CTAs)"
for v in d5 d7 d8 e5 e4; do GOOFY_B200_LIB=$PWD/build/ab/libgoofy_rgb_$v.so python tools/rgb24_ab.py --label "variant $v"; done
done
for r in 2 3 6 8; do GOOFY_B200_RGB24_ROWS_PER_CTA=$r python tools/rgb24_ab.py --label "rows per CTA $r"; done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"encode_rgb24" -f -o gpurun_out/r02_rgb24 python tools/profile_rgb24_target.py > gpurun_out/ncu_rgb24.log 2>&1; echo "ncu rc=$?"
