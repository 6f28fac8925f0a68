#!/bin/bash
# integration/stack/build_examples.sh -- compile the reference's own example programs (sources read in place from
# $REF, nothing copied) against the stack built by build_stack.sh.  Outputs: build/stack/ex/<name>.
#   helloworld: greeter_server / greeter_client   (BASELINE configs[0]: plumbing over TCP)
set -e
ROOT=$(cd "$(dirname "$0")/../.." && pwd)
REF=${REF:-/root/reference}
B=$ROOT/build/stack/grpc
PFX=$ROOT/build/stack/prefix
EX=$ROOT/build/stack/ex
mkdir -p $EX/gen
PROTOC=$B/third_party/protobuf/protoc
$PROTOC -I $REF/examples/protos -I $REF/third_party/protobuf/src --cpp_out=$EX/gen --grpc_out=$EX/gen --plugin=protoc-gen-grpc=$B/grpc_cpp_plugin \
  $REF/examples/protos/helloworld.proto $REF/examples/protos/micro_benchmark.proto
INC="-I$EX/gen -I$REF/include -I$REF -I$REF/third_party/protobuf/src -I$REF/third_party/abseil-cpp -I$PFX/include"
CXXFLAGS="-std=c++14 -O2 -w -include cstdint -DGRPC_USE_IBVERBS"
LIBS="-Wl,--start-group $(find $B -name '*.a' | grep -v -e libprotoc -e plugin_support | tr '\n' ' ') -Wl,--end-group \
  -L$PFX/lib -libverbs -lhdr_histogram -Wl,-rpath,$PFX/lib -lssl -lcrypto -lz -lpthread -ldl -lrt"
for f in helloworld.pb helloworld.grpc.pb micro_benchmark.pb micro_benchmark.grpc.pb; do
  [ $EX/gen/$f.o -nt $EX/gen/$f.cc ] || g++ $CXXFLAGS $INC -c $EX/gen/$f.cc -o $EX/gen/$f.o
done
HW="$EX/gen/helloworld.pb.o $EX/gen/helloworld.grpc.pb.o"
build() { # name source objs...
  local out=$EX/$1 src=$2; shift 2
  g++ $CXXFLAGS $INC $EXTRA -o $out $src "$@" $LIBS
  echo "built $out"
}
build hw_greeter_server $REF/examples/cpp/helloworld/greeter_server.cc $HW
build hw_greeter_client $REF/examples/cpp/helloworld/greeter_client.cc $HW
# client + server of the echo integration test (examples/cpp/test) in one process
g++ $CXXFLAGS $INC -Dmain=server_main -c $REF/examples/cpp/test/greeter_server.cc -o $EX/gen/test_server.o
g++ $CXXFLAGS $INC -Dmain=client_main -c $REF/examples/cpp/test/greeter_client.cc -o $EX/gen/test_client.o
build test_echo_cs $ROOT/integration/stack/cs_main.cc $EX/gen/test_server.o $EX/gen/test_client.o $HW
# the reference's micro-benchmark driver (examples/cpp/micro-bench: the workload of BASELINE configs[1..4]); MPI and
# libnuma are not in the image: single-rank / no-NUMA stand-ins under integration/stack/shim_mb
(cd $B && ninja -j6 absl_flags absl_flags_parse absl_flags_usage absl_flags_usage_internal absl_flags_reflection absl_flags_marshalling \
   absl_flags_internal absl_flags_config absl_flags_program_name absl_flags_commandlineflag absl_flags_commandlineflag_internal \
   absl_flags_private_handle_accessor > /dev/null)
LIBS="-Wl,--start-group $(find $B -name '*.a' | grep -v -e libprotoc -e plugin_support | tr '\n' ' ') -Wl,--end-group \
  -L$PFX/lib -libverbs -lhdr_histogram -Wl,-rpath,$PFX/lib -lssl -lcrypto -lz -lpthread -ldl -lrt"
MB="$EX/gen/micro_benchmark.pb.o $EX/gen/micro_benchmark.grpc.pb.o"
EXTRA="-I$ROOT/integration/stack/shim_mb" build mb_server $REF/examples/cpp/micro-bench/mb_server.cc $MB
EXTRA="-I$ROOT/integration/stack/shim_mb -I$REF/examples/cpp/micro-bench" build mb_client $REF/examples/cpp/micro-bench/mb_client.cc $MB
