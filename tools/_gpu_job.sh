set -x
mkdir -p gpurun_out/r2a
python -m pytest tests -m gpu -q 2>&1 | tail -25 > gpurun_out/r2a/pytest.log
python tools/refinit_probe.py fp32 fp16 bf16 > gpurun_out/r2a/probe_default.jsonl 2> gpurun_out/r2a/probe_default.err
M0_TC_FUSE_SE=0 python tools/refinit_probe.py fp16 bf16 > gpurun_out/r2a/probe_unfused.jsonl 2>> gpurun_out/r2a/probe_default.err
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2a/smoke.log 2>&1
python bench.py --steps 20 --warmup 5 > gpurun_out/r2a/bench.json 2> gpurun_out/r2a/bench.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'search_select|search_expand|search_begin' -c 6 -o gpurun_out/r2a/tree python bench.py --steps 12 --warmup 3 > gpurun_out/r2a/ncu_tree.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'attention_tc|gemm_tc_kernel|ln_res_gn|se_hidden|gn_act_res' -c 10 -o gpurun_out/r2a/attn python bench.py --steps 2 --warmup 1 > gpurun_out/r2a/ncu_attn.log 2>&1
ls -la gpurun_out/r2a
