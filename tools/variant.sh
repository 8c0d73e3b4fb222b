#!/bin/bash
# tools/variant.sh NAME [nvcc -D flags...] — tuning build of the predict kernel: build/variants/libvmis_NAME.so
NAME=$1; shift
cd "$(dirname "$0")/../serenade_b200/csrc"
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC,-O3,-pthread -Xptxas -v "$@" -c -o ../../build/obj/predict_$NAME.o predict_sm100.cu 2>&1 | grep -A2 "vmis_predict_kernel" | grep -E "registers|spill" | tr '\n' ' '
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../../build/variants/libvmis_$NAME.so ../../build/obj/predict_$NAME.o $(ls ../../build/obj/*.o | grep -v "/predict_") -lz && echo " -> build/variants/libvmis_$NAME.so"
