#!/bin/sh
# Generates the "ref + 64-bit patch" variant of the reference's svo.cu into a
# git-ignored temporary file.  Five edits, all confined to key arithmetic that truncates to 32 bits:
#  1. depthFromKey (svo.cu:68-78): table lookup valid for keys < 2^31  ->  (63 - clzll(key)) / 3
#  2. getFirstValueAndShiftDown (svo.cu:87-88): int shifts -> octkey shifts
#  3. struct negative (svo.cu:174): operator()(const int)  -> (const octkey)
#  4. struct depth_is_zero (svo.cu:444): operator()(const int) -> (const octkey)
#  5. getOccupiedChildren (svo.cu:524): int child_val -> octkey child_val
set -e
in="$1"; out="$2"
sed \
  -e '/^__device__ int depthFromKey(octkey key) {/,/^}/c\__device__ int depthFromKey(octkey key) { return (63 - __clzll(key)) / 3; }' \
  -e 's/key -= ((8 + value) << 3 \* (depth - 1));/key -= (((octkey)(8 + value)) << 3 * (depth - 1));/' \
  -e 's/key += (1 << 3 \* (depth - 1));/key += (((octkey)1) << 3 * (depth - 1));/' \
  -e 's/__host__ __device__ bool operator() (const int x) {/__host__ __device__ bool operator() (const octkey x) {/' \
  -e 's/__device__ bool operator() (const int key) {/__device__ bool operator() (const octkey key) {/' \
  -e 's/    int child_val = -1;/    octkey child_val = -1;/' \
  "$in" > "$out"
# sanity: all five edits must have landed
grep -q '__clzll' "$out"
test "$(grep -c '(octkey)' "$out")" -ge 2
grep -q 'operator() (const octkey x)' "$out"
grep -q 'operator() (const octkey key)' "$out"
grep -q 'octkey child_val = -1;' "$out"
