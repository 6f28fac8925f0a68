#!/bin/bash
# integration/stack/build_stack.sh -- build the reference's full gRPC stack OFFLINE, out of tree, into
# build/stack/ (git-ignored, gpurun-ignored).  Two un-vendored dependencies are stood in for:
#   libibverbs      -> oracle/shim/fake_verbs.cc (in-process loopback verbs; TEST INFRASTRUCTURE)
#   HdrHistogram_c  -> integration/stack/hdr_histogram_mini.c (own minimal implementation)
# zlib and OpenSSL come from the image (the vendored zlib's CMake renames a file in the read-only tree).
# Usage: build_stack.sh [ninja targets...]   (default: grpc++ grpc_cpp_plugin protoc)
set -e
ROOT=$(cd "$(dirname "$0")/../.." && pwd)
REF=${REF:-/root/reference}
OUT=$ROOT/build/stack
PFX=$OUT/prefix
mkdir -p $PFX/include/infiniband $PFX/include/hdr $PFX/lib $OUT/grpc
cp $ROOT/oracle/shim/infiniband/verbs.h $PFX/include/infiniband/verbs_base.h
cp $ROOT/integration/stack/infiniband/verbs.h $PFX/include/infiniband/verbs.h
cp $ROOT/integration/stack/hdr/hdr_histogram.h $PFX/include/hdr/hdr_histogram.h
g++ -O2 -g -fPIC -shared -std=c++14 -I$PFX/include -o $PFX/lib/libibverbs.so $ROOT/oracle/shim/fake_verbs.cc $ROOT/integration/stack/verbs_event_stubs.cc -lpthread
gcc -O2 -g -fPIC -shared -I$PFX/include -o $PFX/lib/libhdr_histogram.so $ROOT/integration/stack/hdr_histogram_mini.c
export CPLUS_INCLUDE_PATH=$PFX/include C_INCLUDE_PATH=$PFX/include  # grpc++ sources reach infiniband/verbs.h through core headers
cd $OUT/grpc
if [ ! -f build.ninja ]; then
  cmake -G Ninja $REF -DCMAKE_POLICY_VERSION_MINIMUM=3.5 -DCMAKE_BUILD_TYPE=Release \
    -DCMAKE_CXX_FLAGS="-include cstdint -include array -include algorithm -include mutex -include thread -include chrono -include string -include condition_variable -include functional -include limits -w" -DCMAKE_C_FLAGS="-w" \
    -DgRPC_BUILD_TESTS=OFF -DgRPC_BUILD_CSHARP_EXT=OFF -DgRPC_INSTALL=OFF \
    -DgRPC_BUILD_GRPC_CSHARP_PLUGIN=OFF -DgRPC_BUILD_GRPC_NODE_PLUGIN=OFF -DgRPC_BUILD_GRPC_OBJECTIVE_C_PLUGIN=OFF \
    -DgRPC_BUILD_GRPC_PHP_PLUGIN=OFF -DgRPC_BUILD_GRPC_PYTHON_PLUGIN=OFF -DgRPC_BUILD_GRPC_RUBY_PLUGIN=OFF \
    -DgRPC_ZLIB_PROVIDER=package -DgRPC_SSL_PROVIDER=package \
    -DIBVERBS_ROOT_DIR=$PFX -DHdrHistogram_ROOT_DIR=$PFX/include \
    -DCMAKE_LIBRARY_PATH=$PFX/lib -DCMAKE_INCLUDE_PATH=$PFX/include
fi
TARGETS=${@:-grpc++ grpc++_reflection grpc_cpp_plugin protoc}
ninja -j${JOBS:-6} $TARGETS
