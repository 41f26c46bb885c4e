#!/bin/bash
# sm_100a SASS listings of the kernels north_star names, from the in-tree objects (no GPU needed): profiles/sass/
set -e
cd "$(dirname "$0")/.."
L=biceps_b200/lib; O=profiles/sass
dump() { obj=$1; pat=$2; out=$3
  fn=$(cuobjdump -elf $L/$obj 2>/dev/null | grep -o "\.text\.[A-Za-z0-9_]*" | sed 's/^\.text\.//' | sort -u | grep -E "$pat" | head -1)
  [ -n "$fn" ] || { echo "no kernel matching $pat in $obj"; return; }
  cuobjdump -sass -fun "$fn" $L/$obj | sed 's/ *\/\* 0x[0-9a-f]* \*\/$//' > $O/$out
  echo "$out  <- $fn  ($(grep -c '^ *\/\*[0-9a-f]\{4,6\}\*\/' $O/$out) instructions)"; }
dump sampler_spec.o 'mcmc_spec_kernelILi4ELi4ELb1ELb0' mcmc_spec_kernel_4_4_gamma.sass
dump sampler_tab.o 'mcmc_tab_kernelILi4ELb1' mcmc_tab_kernel_4_gamma.sass
dump sampler_ws.o 'mcmc_ws_kernelILi4ELi4ELb1ELi' mcmc_ws_kernel_4_4_gamma.sass
dump sampler_fast.o 'mcmc_fast_kernelILi4ELb1ELb1' mcmc_fast_kernel_4_gamma_mk.sass
dump sampler_pf2.o 'mcmc_pf2_kernelILi1ELi4' mcmc_pf2_kernel_1_4.sass
dump sampler_pf2.o 'mcmc_pf2_kernelILi3ELi4' mcmc_pf2_kernel_3_4.sass
dump sampler_pf.o 'mcmc_pf_kernel' mcmc_pf_kernel.sass
dump mbar.o 'mbar_pass_reg_kernelILi11ELb1ELi3' mbar_pass_reg_kernel_11_hess.sass
dump mbar.o 'mbar_pass_reg_kernelILi11ELb0ELi2' mbar_pass_reg_kernel_11_plain.sass
dump mbar.o 'mbar_pop_reg_kernelILi11' mbar_pop_reg_kernel_11.sass
dump mbar.o '10ukn_kernel' ukn_kernel.sass
dump sampler.o 'ukn_traj_kernel' ukn_traj_kernel.sass
dump precompute.o 'row_kernelILi0ELi0' row_kernel_sse.sass
dump precompute.o 'row_kernelILi2ELi1' row_kernel_gamma_expref.sass
dump precompute.o 'col_stats_kernelILb1' col_stats_kernel_var.sass
dump precompute.o 'logz_kernel' logz_kernel.sass
dump convergence.o 'autocorr_partial_kernelILi16' autocorr_partial_kernel_16.sass
dump convergence.o 'jsd_kernel' jsd_kernel.sass
