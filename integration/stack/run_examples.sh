#!/bin/bash
# integration/stack/run_examples.sh -- run the reference's own examples on the offline-built reference stack (CPU):
#   1. helloworld over TCP (BASELINE configs[0]: "Hello world")
#   2. examples/cpp/test echo (1000 random messages of 1 B .. 4 MiB - 1 KiB, GPR_ASSERT(msg == reply)) over TCP
#   3. the same echo with GRPC_PLATFORM_TYPE=RDMA_BPEV: the reference's own endpoint, pair, ring buffer and
#      poller over the in-process loopback verbs
# Prints one line per run; needs build_stack.sh + build_examples.sh.
ROOT=$(cd "$(dirname "$0")/../.." && pwd)
EX=$ROOT/build/stack/ex
cd $EX
GRPC_PLATFORM_TYPE=TCP GRPC_VERBOSITY=INFO ./hw_greeter_server 50071 > hw_srv.log 2>&1 &
SP=$!
sleep 1.5
OUT=$(GRPC_PLATFORM_TYPE=TCP timeout 20 ./hw_greeter_client 2>/dev/null | tail -1)
kill $SP 2>/dev/null; wait $SP 2>/dev/null
echo "helloworld TCP: '$OUT' ($(grep -c 'Select TCP mode' hw_srv.log) x 'Select TCP mode' in the server log)"
for MODE in TCP RDMA_BPEV; do
  T0=$(date +%s.%N)
  GRPC_PLATFORM_TYPE=$MODE GRPC_VERBOSITY=INFO GRPC_RDMA_RING_BUFFER_SIZE_KB=16384 timeout 600 ./test_echo_cs 5008$((RANDOM % 10)) > echo_$MODE.log 2>&1
  RC=$?
  T1=$(date +%s.%N)
  echo "examples/cpp/test echo $MODE: rc=$RC, $(grep -c 'received\.' echo_$MODE.log) replies equal to their request, mode lines: $(grep -o 'Select [A-Za-z ]* mode' echo_$MODE.log | sort | uniq -c | tr '\n' ';'), $(python3 -c "print(round($T1 - $T0, 1))") s"
done
# the reference's micro-benchmark driver (unary, 1 KiB requests) over TCP: server and client as two processes
GRPC_PLATFORM_TYPE=TCP timeout 40 ./mb_server --port=50079 --threads=2 --cqs=2 --resp=8 > mb_srv.log 2>&1 &
SP=$!
sleep 1.5
GRPC_PLATFORM_TYPE=TCP timeout 20 stdbuf -oL ./mb_client --target=localhost:50079 --req=1024 --concurrent=1 --duration=3 --warmup=100 > mb_cli.log 2>&1
echo "micro-bench TCP (mb_server + mb_client): rc=$?, $(grep Aggregated mb_cli.log), $(tail -1 mb_cli.log)"
kill $SP 2>/dev/null; wait $SP 2>/dev/null
