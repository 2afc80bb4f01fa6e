#!/bin/bash
# Host / PCIe / NUMA topology of the GPU box (for the multi-GPU upload path): one text file.
mkdir -p gpurun_out
{
  echo "== nvidia-smi topo -m"; nvidia-smi topo -m
  echo "== lscpu"; lscpu | grep -E "^CPU\(s\)|Model name|Socket|NUMA|Thread|Core"
  echo "== affinity of this shell"; taskset -p $$ 2>/dev/null; grep -E "Cpus_allowed_list|Mems_allowed_list" /proc/self/status
  echo "== GPUs: pci bus id -> numa node"
  for d in $(nvidia-smi --query-gpu=pci.bus_id --format=csv,noheader); do
    b=$(echo ${d#0000} | tr 'A-Z' 'a-z'); echo "$d numa_node=$(cat /sys/bus/pci/devices/$b/numa_node 2>/dev/null) local_cpulist=$(cat /sys/bus/pci/devices/$b/local_cpulist 2>/dev/null)"
  done
  echo "== nodes"; for n in /sys/devices/system/node/node*; do echo "$n cpulist=$(cat $n/cpulist) $(grep MemFree $n/meminfo)"; done
  echo "== memory"; grep -E "MemTotal|MemAvailable" /proc/meminfo; nproc
} > gpurun_out/topo.txt 2>&1
tail -n 30 gpurun_out/topo.txt
