#!/bin/bash
# developer helper: build _probe/libtb_<name>.so = the library with fast.cu/frontend.cu (or the files named in FILES) recompiled with extra -D flags
# usage: scripts/build_variant.sh NAME -DFOO=1 ...   (run after the normal build; TB_LIB_PATH=_probe/libtb_NAME.so selects it)
set -e
name=$1; shift
cd "$(dirname "$0")/.."
B=trackingbench_slam_b200/build; C=trackingbench_slam_b200/csrc
mkdir -p _probe/$name
objs=""
for f in $B/*.o; do objs="$objs $f"; done
for src in ${FILES:-fast frontend}; do
  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -Xcompiler -O2 --fmad=false "$@" -c -o _probe/$name/$src.o $C/$src.cu &
  objs=$(echo $objs | sed "s#$B/$src.o#_probe/$name/$src.o#")
done
wait
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o _probe/libtb_$name.so $objs
echo built _probe/libtb_$name.so
