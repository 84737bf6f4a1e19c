#!/bin/bash
# First GPU job of round 2: is the reference's third-party FNO implementation importable on the
# GPU box?  Output is committed as profiles/r02_neuralop_probe.txt.
echo "== host: $(hostname)  date: $(date -u +%FT%TZ)"
echo "== python -c 'import neuralop, tltorch, tensorly'"
python -c "import neuralop, tltorch, tensorly; print(neuralop.__file__)" 2>&1 | tail -3
for m in neuralop tltorch tensorly opt_einsum torch_harmonics timm; do
  python -c "import $m; print('$m', getattr($m,'__version__','?'), $m.__file__)" 2>&1 | tail -1
done
echo "== pip list | grep -i -E 'neural|tensorly|tltorch|opt.einsum|harmonics|timm'"
python -m pip list 2>/dev/null | grep -i -E "neural|tensorly|tltorch|opt.einsum|harmonics|timm" || echo "(no match)"
echo "== ls baseline/_ref"
ls -la baseline/_ref 2>&1 | head
echo "== ls /opt/wheelhouse | grep -i -E 'neural|tensorly|tltorch'"
ls /opt/wheelhouse 2>/dev/null | grep -i -E "neural|tensorly|tltorch" || echo "(no match)"
echo "== /root/reference present on this box?"
ls -d /root/reference 2>&1
