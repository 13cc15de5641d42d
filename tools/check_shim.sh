#!/bin/bash
# Type-checks js/napi/shim.cc against include/rfx.h and a stub of the N-API prototypes it uses (there is no Node in this image).
cd "$(dirname "$0")/.."
g++ -std=c++17 -fsyntax-only -Wall -Itools/napi_stub -Iinclude js/napi/shim.cc && echo "shim.cc: OK (type-checked against include/rfx.h)"
