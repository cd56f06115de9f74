#!/usr/bin/env bash
# Install the UNMODIFIED reference (antoine311200/sow, /root/reference) into baseline/_ref (git-ignored; travels to the
# GPU box with the gpurun snapshot).  Build-container only: /root/reference does not exist on the GPU box.
#
# The reference's setup.py lists packages=["tn_gradient"] only, so the stock wheel drops tn_gradient/layer and
# tn_gradient/optimizer (namespace sub-packages without __init__.py) and the installed package cannot even import
# tn_gradient.prepare.  The install therefore runs from a /tmp copy whose setup.py names the three package
# directories; every .py file of the package is installed byte-identical (sha256 checked below).
set -euo pipefail
HERE="$(cd "$(dirname "$0")" && pwd)"
REF="${SOW_REFERENCE_ROOT:-/root/reference}"
TMP="$(mktemp -d)"
cp -r "$REF" "$TMP/ref"
sed -i 's/packages=\["tn_gradient"\]/packages=["tn_gradient", "tn_gradient.layer", "tn_gradient.optimizer"]/' "$TMP/ref/setup.py"
rm -rf "$HERE/_ref"
python -m pip install --no-index --no-build-isolation --no-deps --find-links /opt/wheelhouse \
    --target "$HERE/_ref" "$TMP/ref" >/dev/null
find "$HERE/_ref" -name __pycache__ -type d -exec rm -rf {} +
( cd "$REF" && find tn_gradient -name '*.py' | sort | xargs sha256sum ) > "$TMP/a.sha"
( cd "$HERE/_ref" && find tn_gradient -name '*.py' | sort | xargs sha256sum ) > "$TMP/b.sha"
diff "$TMP/a.sha" "$TMP/b.sha" && echo "baseline/_ref: $(wc -l < "$TMP/b.sha") files, byte-identical to $REF/tn_gradient"
rm -rf "$TMP"
