for w in graph_coloring job_shop; do
  for c in 6 8 12 16; do SFGPU_SPEC_CHUNKS=$c python scripts/kbench.py $w 2>&1 | tail -1; done
  KB_MATERIALISE=0 python scripts/kbench.py $w 2>&1 | tail -1
done
python -m pytest tests/test_gpu_spec.py -x -q 2>&1 | tail -3
