# debug variant of the library with the CTA-0 event trace compiled into the flat conv kernel -> tools/libcilrs_trace.so
set -e
cd "$(dirname "$0")/.."
S=cilrs-autonomous-driving-carla_b200/csrc
mkdir -p /tmp/cilrs_trace
for f in api conv conv_flat model preprocess fp32_path jpeg pipeline; do
  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -DCF_TRACE $EXTRA -c $S/$f.cu -o /tmp/cilrs_trace/$f.o &
done
wait
nvcc -shared -o tools/libcilrs_trace.so /tmp/cilrs_trace/*.o -gencode arch=compute_100a,code=sm_100a -cudart static
ls -la tools/libcilrs_trace.so
