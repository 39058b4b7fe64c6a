#!/usr/bin/env bash
# Builds the host-side C++ library (model, parser, presolve, phase control, dense tableau
# helpers) and the `simplex` CLI, linked against the CUDA C-ABI library.
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
PKG="${HERE}/.."
CXX=g++
LIBSRCS=$(ls "${HERE}"/*.cpp | grep -v '/main.cpp$' || true)
${CXX} -std=c++17 -O2 -ffp-contract=off -fPIC -shared -Wall -I"${PKG}/../include" \
    -o "${PKG}/libsimplex_b200_host.so" ${LIBSRCS} -L"${PKG}" -lsimplex_b200 -lpthread -Wl,-rpath,'$ORIGIN'
if [ -f "${HERE}/main.cpp" ]; then
    ${CXX} -std=c++17 -O2 -Wall -I"${PKG}/../include" -o "${PKG}/simplex" "${HERE}/main.cpp" \
        -L"${PKG}" -lsimplex_b200_host -lsimplex_b200 -Wl,-rpath,'$ORIGIN'
fi
${CXX} -std=c++17 -O1 -Wall -I"${PKG}/../include" -o "${PKG}/host_unit_tests" \
    "${HERE}/selftest/host_unit_tests.cpp" -L"${PKG}" -lsimplex_b200_host -lsimplex_b200 -Wl,-rpath,'$ORIGIN'
echo "built ${PKG}/libsimplex_b200_host.so, ${PKG}/simplex, ${PKG}/host_unit_tests"
