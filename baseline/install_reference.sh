#!/bin/sh
# Install the UNMODIFIED reference into baseline/_ref (git-ignored; travels to the GPU box).
# The build writes into the source tree, /root/reference is read-only -> install from a scratch copy.
# --no-deps: biopython / python-nexus / newick / pulp / matplotlib are not in the offline wheelhouse and
# are not on the numeric path (baseline/ref_loader.py registers inert stand-ins for the two that are
# imported at module scope).
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
SRC="${1:-/root/reference}"
TMP="$(mktemp -d /tmp/refsrc.XXXXXX)"
cp -r "$SRC"/. "$TMP"/ && chmod -R u+w "$TMP"
rm -rf "$HERE/_ref"
python -m pip install --no-index --no-build-isolation --no-deps --find-links /opt/wheelhouse \
    --target "$HERE/_ref" "$TMP"
rm -rf "$TMP"
