#!/bin/bash
# round 2, session W: packed-RGB kernels: widening on the multiply pipe (IMAD.HI + IMAD) vs byte permutes
for rep in 1 2; do
python tools/rgb24_ab.py --label "product (PRMT widen)"
GOOFY_B200_LIB=$PWD/build/ab/libgoofy_rgb_mul.so python tools/rgb24_ab.py --label "IMAD widen"
GOOFY_B200_LIB=$PWD/build/ab/libgoofy_rgb_mul5.so python tools/rgb24_ab.py --label "IMAD widen, dual 5 CTAs"
GOOFY_B200_LIB=$PWD/build/ab/libgoofy_rgb_mul5.so GOOFY_B200_RGB24_ROWS_PER_CTA=8 python tools/rgb24_ab.py --label "IMAD widen, dual 5 CTAs, 8 rows"
GOOFY_B200_LIB=$PWD/build/ab/libgoofy_rgb_mul5.so GOOFY_B200_RGB24_ROWS_PER_CTA=6 python tools/rgb24_ab.py --label "IMAD widen, dual 5 CTAs, 6 rows"
done
