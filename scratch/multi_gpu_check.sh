#!/bin/bash
# byte-identity of the CLI outputs between one process and torchrun with N ranks (NCCL), HDF5 and dense inputs
set -e
N=${1:-2}
G=tests/golden
O=gpurun_out/mg
rm -rf $O; mkdir -p $O
python -m yanocomp_b200 pack $G/hek293t_wt.collapsed.hdf5 $O/wt.ycd
python -m yanocomp_b200 pack $G/hek293t_mettl3.collapsed.hdf5 $O/mettl3.ycd
python -m yanocomp_b200 gmmtest -c $G/hek293t_wt.collapsed.hdf5 -t $G/hek293t_mettl3.collapsed.hdf5 -o $O/h5_1.bed -s $O/h5_1.json --random-seed 0
python -m yanocomp_b200 gmmtest -c $O/wt.ycd -t $O/mettl3.ycd -o $O/ycd_1.bed -s $O/ycd_1.json --random-seed 0
python -m yanocomp_b200 gmmtest -c $O/wt.ycd -t $O/mettl3.ycd -o $O/ycd_1_nosm.bed --random-seed 0
torchrun --standalone --local-addr 127.0.0.1 --nproc-per-node $N -m yanocomp_b200 gmmtest -c $G/hek293t_wt.collapsed.hdf5 -t $G/hek293t_mettl3.collapsed.hdf5 -o $O/h5_N.bed -s $O/h5_N.json --random-seed 0
torchrun --standalone --local-addr 127.0.0.1 --nproc-per-node $N -m yanocomp_b200 gmmtest -c $O/wt.ycd -t $O/mettl3.ycd -o $O/ycd_N.bed -s $O/ycd_N.json --random-seed 0
torchrun --standalone --local-addr 127.0.0.1 --nproc-per-node $N -m yanocomp_b200 gmmtest -c $O/wt.ycd -t $O/mettl3.ycd -o $O/ycd_N_nosm.bed --random-seed 0
sed -i 's#gpurun_out/mg/[a-z0-9]*\.ycd#X#g; s#tests/golden/[a-z0-9_.]*hdf5#X#g' $O/*.json
md5sum $O/*.bed $O/*.json
wc -l $O/h5_1.bed
cmp $O/h5_1.bed $O/h5_N.bed && cmp $O/h5_1.bed $O/ycd_1.bed && cmp $O/h5_1.bed $O/ycd_N.bed && cmp $O/h5_1.bed $O/ycd_N_nosm.bed && cmp $O/h5_1.bed $O/ycd_1_nosm.bed && cmp $O/h5_1.json $O/h5_N.json && cmp $O/ycd_1.json $O/ycd_N.json && cmp $O/h5_1.json $O/ycd_1.json && echo "ALL IDENTICAL over 1 and $N ranks"
