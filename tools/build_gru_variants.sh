# Builds libnpd_<name>.so variants of the GRU pair kernel for tools/exp_gru_precision.py (round-2 precision experiments,
# DESIGN.md 4.3e).  Needs the regular objects (make -C neural_polar_decoder_b200/csrc) to exist.
#   bash tools/build_gru_variants.sh w3 "-DNPD_GRU_ZC=0 -DNPD_GRU_LO=0"  w1 "-DNPD_GRU_ZC=1 -DNPD_GRU_LO=0"  a6 "-DNPD_GRU_ACT=6"
set -e
cd "$(dirname "$0")/../neural_polar_decoder_b200/csrc"
NV="/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -I../../include -std=c++17 -O3 -lineinfo -Xcompiler -fPIC,-fvisibility=hidden -DNPD_HAVE_GRU -DNPD_HAVE_CONV -Xptxas -v"
while [ $# -ge 2 ]; do
  name=$1; flags=$2; shift 2
  $NV $flags -c gru_decode.cu -o /tmp/gru_$name.o 2> /tmp/gru_$name.log
  grep -A2 "gru_decode_kernel3ILi8" /tmp/gru_$name.log | grep -E "spill|Used" | tr '\n' ' '; echo " <- $name ($flags)"
  /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../libnpd_$name.so npd_api.o sc_decode.o scl_decode.o \
      encode_channel.o count_sweep.o /tmp/gru_$name.o conv_net.o host_pipe.o gru_train.o -L/usr/local/cuda/lib64 -lcudart -lcublas \
      -Xlinker -rpath=/usr/local/cuda/lib64 -lpthread -ldl -lrt
done
