{
  # Drop-in for the reference's binding.gyp (target name and output path unchanged:
  # build/Release/cpp_utils.node, required by cpp_utils/cpp_utils.js:4).  The addon is the
  # N-API shim only; the CUDA code lives in libycnr_als.so built by nvcc
  # (python -m you_can_not_recommend_b200.build) and is linked, not compiled, here.
  "targets": [
    {
      "target_name": "cpp_utils",
      "sources": [ "ycnr_napi.cc" ],
      "include_dirs": [ "../../include" ],
      "defines": [ "NAPI_VERSION=6" ],
      "libraries": [
        "-L<(module_root_dir)/..", "-lycnr_als",
        "-Wl,-rpath,<(module_root_dir)/.."
      ],
      "cflags_cc": [ "-std=c++14", "-O2" ]
    }
  ]
}
