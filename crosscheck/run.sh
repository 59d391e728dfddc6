#!/bin/bash
# usage: crosscheck/run.sh <sigma0-polymath checkout> [YYYY-MM-DD]
# Builds the UNMODIFIED reference against arkworks revisions resolved by date and writes the vectors that
# tests/test_crosscheck_cpu.py compares with the oracle.  Needs cargo, git and network access.
set -euo pipefail
here="$(cd "$(dirname "$0")" && pwd)"
ref="${1:?path of a sigma0-dev/polymath checkout}"
when="${2:-}"
if [ -z "$when" ]; then
  when="$(git -C "$ref" log -1 --format=%cs 2>/dev/null || true)"
fi
[ -n "$when" ] || { echo "cannot tell the date of the reference checkout: pass YYYY-MM-DD" >&2; exit 2; }
ln -sfn "$(cd "$ref" && pwd)" "$here/reference"
work="$(mktemp -d)"
trap 'rm -rf "$work"' EXIT
resolve() {   # <url> <branch> -> sha of the last commit on <branch> not after $when
  local url="$1" branch="$2" dir="$work/$(basename "$1")"
  [ -d "$dir" ] || git clone -q --filter=blob:none --no-checkout -b "$branch" "$url" "$dir"
  git -C "$dir" rev-list -1 --before="$when 23:59:59" "$branch"
}
alg="$(resolve https://github.com/arkworks-rs/algebra master)"
r1cs="$(resolve https://github.com/arkworks-rs/r1cs-std master)"
cp_="$(resolve https://github.com/arkworks-rs/crypto-primitives main)"
snark="$(resolve https://github.com/arkworks-rs/snark master)"
sed -i -E \
  -e "s#(arkworks-rs/algebra/\", )(branch|rev) = \"[^\"]*\"#\1rev = \"$alg\"#" \
  -e "s#(arkworks-rs/r1cs-std/\", )(branch|rev) = \"[^\"]*\"#\1rev = \"$r1cs\"#" \
  -e "s#(arkworks-rs/crypto-primitives/\", )(branch|rev) = \"[^\"]*\"#\1rev = \"$cp_\"#" \
  -e "s#(arkworks-rs/snark/\", )(branch|rev) = \"[^\"]*\"#\1rev = \"$snark\"#" \
  "$here/Cargo.toml"
out="$here/../tests/golden/from_reference"
mkdir -p "$out"
(cd "$here" && cargo run --release -- "$out" "$here/../tests/golden/kernels.json")
cp "$here/Cargo.lock" "$out/Cargo.lock"
cat > "$out/PROVENANCE.json" <<JSON
{"reference_commit": "$(git -C "$ref" rev-parse HEAD 2>/dev/null || echo unknown)", "date": "$when",
 "algebra": "$alg", "r1cs_std": "$r1cs", "crypto_primitives": "$cp_", "snark": "$snark",
 "rustc": "$(rustc --version)"}
JSON
echo "vectors written to $out; now run: python -m pytest tests/test_crosscheck_cpu.py -q"
