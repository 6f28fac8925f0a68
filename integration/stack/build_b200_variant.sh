#!/bin/bash
# integration/stack/build_b200_variant.sh -- the reference's gRPC core with the B200 pair library swapped in.
# The reference's OWN endpoint (rdma_bp_posix.cc) and event engines (ev_epollex_rdma_bp[ev]_linux.cc,
# ev_epoll1_rdma_bpev_linux.cc) are recompiled UNCHANGED, in place, with integration/shim/ in front of the
# include path (so grpc_core::ibverbs::{PairPollable,PairPool,Poller} forward to include/b200_pair.h), put into a
# copy of libgrpc.a instead of the originals, and the example programs are relinked against it and
# libb200rdma.so: grpc_rdma_bp_create (rdma_bp_posix.cc:706) then runs over HBM rings and sm_100a kernels.
# Needs build_stack.sh + build_examples.sh first.  Outputs (git-ignored, travel to the GPU box): integration/_bin/.
set -e
ROOT=$(cd "$(dirname "$0")/../.." && pwd)
REF=${REF:-/root/reference}
B=$ROOT/build/stack/grpc
PFX=$ROOT/build/stack/prefix
EX=$ROOT/build/stack/ex
V=$ROOT/build/stack/b200
OUT=$ROOT/integration/_bin
mkdir -p $V $OUT
INC="-I$ROOT/integration/shim -I$ROOT/include -I$REF/third_party/re2 -I$REF/include -I$REF -I$REF/third_party/address_sorting/include \
 -I$PFX/include -I$B/third_party/re2 -I$REF/src/core/ext/upb-generated -I$REF/src/core/ext/upbdefs-generated -I$REF/third_party/upb \
 -I$REF/third_party/xxhash -I$B/third_party/cares/cares -I$REF/third_party/cares/cares -I$REF/third_party/abseil-cpp"
FLAGS="-DCARES_STATICLIB -DGRPC_USE_IBVERBS -DHAVE_IBVERBS -include cstdint -include array -include algorithm -include mutex \
 -include thread -include chrono -include string -include condition_variable -include functional -include limits -w -g -O3 -DNDEBUG -std=c++14 -fPIC"
OBJS=""
for f in rdma_bp_posix ev_epollex_rdma_bp_linux ev_epollex_rdma_bpev_linux ev_epoll1_rdma_bpev_linux; do
  g++ $FLAGS $INC -c $REF/src/core/lib/iomgr/$f.cc -o $V/$f.cc.o
  OBJS="$OBJS $V/$f.cc.o"
done
cp $B/libgrpc.a $V/libgrpc_b200.a
ar d $V/libgrpc_b200.a rdma_bp_posix.cc.o ev_epollex_rdma_bp_linux.cc.o ev_epollex_rdma_bpev_linux.cc.o ev_epoll1_rdma_bpev_linux.cc.o
ar r $V/libgrpc_b200.a $OBJS
# of the reference's ibverbs library only Config (env parsing) is still needed
(cd $V && ar x $B/libgrpc_ibverbs.a config.cc.o)
LIBS="-Wl,--start-group $V/libgrpc_b200.a $V/config.cc.o $(find $B -name '*.a' | grep -v -e libprotoc -e plugin_support -e 'libgrpc\.a' -e libgrpc_ibverbs | tr '\n' ' ') -Wl,--end-group \
  -L$ROOT/grpc-rdma_b200/lib -lb200rdma -L$PFX/lib -libverbs -lhdr_histogram -lssl -lcrypto -lz -lpthread -ldl -lrt \
  -Wl,-rpath,\$ORIGIN -Wl,-rpath,\$ORIGIN/../../grpc-rdma_b200/lib"
EXINC="-I$EX/gen -I$REF/include -I$REF -I$REF/third_party/protobuf/src -I$REF/third_party/abseil-cpp -I$PFX/include"
HW="$EX/gen/helloworld.pb.o $EX/gen/helloworld.grpc.pb.o"
g++ -std=c++14 -O2 -w -include cstdint $EXINC -o $OUT/test_echo_cs_b200 $ROOT/integration/stack/cs_main.cc $EX/gen/test_server.o $EX/gen/test_client.o $HW $LIBS
# the in-process fake verbs + the mini HdrHistogram are still linked (RDMA_EVENT mode objects reference them): ship them beside the binary
cp $PFX/lib/libibverbs.so $PFX/lib/libhdr_histogram.so $OUT/
echo "built $OUT/test_echo_cs_b200"; ls -la $OUT
# the reference's micro-benchmark driver over the same swapped core (two processes: they meet on the CUDA-IPC wire)
MB="$EX/gen/micro_benchmark.pb.o $EX/gen/micro_benchmark.grpc.pb.o"
MBINC="$EXINC -I$ROOT/integration/stack/shim_mb -I$REF/examples/cpp/micro-bench"
g++ -std=c++14 -O2 -w -include cstdint -DGRPC_USE_IBVERBS $MBINC -o $OUT/mb_server_b200 $REF/examples/cpp/micro-bench/mb_server.cc $MB $LIBS
g++ -std=c++14 -O2 -w -include cstdint -DGRPC_USE_IBVERBS $MBINC -o $OUT/mb_client_b200 $REF/examples/cpp/micro-bench/mb_client.cc $MB $LIBS
strip $OUT/test_echo_cs_b200 $OUT/mb_server_b200 $OUT/mb_client_b200
ls -la $OUT
