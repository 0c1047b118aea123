#!/bin/bash
# What sits between the host memory and each GPU on this box: NUMA nodes the container may use, the PCIe path of every
# GPU (shared switches / root ports), link generation and width.  Read together with tools/micro/pcie_multi.cu.
echo "== cpus / numa"
nproc; lscpu 2>/dev/null | grep -i -E "model name|socket|numa|^cpu\(s\)|thread"
grep -E "Cpus_allowed_list|Mems_allowed_list" /proc/self/status
for n in /sys/devices/system/node/node*; do [ -d "$n" ] && echo "$(basename $n): cpus $(cat $n/cpulist) $(grep MemTotal $n/meminfo | awk '{print $4/1048576 " GiB"}')"; done
free -g | head -2
echo "== nvidia-smi topo"
nvidia-smi topo -m 2>/dev/null | sed 's/\x1b\[[0-9;]*m//g'
echo "== per GPU"
nvidia-smi --query-gpu=index,pci.bus_id,pcie.link.gen.current,pcie.link.gen.max,pcie.link.width.current,pcie.link.width.max --format=csv,noheader 2>/dev/null
for id in $(nvidia-smi --query-gpu=pci.bus_id --format=csv,noheader 2>/dev/null); do
    d=$(echo "$id" | tr 'A-Z' 'a-z' | sed 's/^0000//')
    p=/sys/bus/pci/devices/$d
    [ -e "$p" ] || continue
    echo "$d numa_node=$(cat $p/numa_node 2>/dev/null) local_cpulist=$(cat $p/local_cpulist 2>/dev/null) path=$(readlink -f $p | sed 's,/sys/devices/,,')"
done
echo "== lspci tree"
lspci -tv 2>/dev/null | head -80 || echo "no lspci"
