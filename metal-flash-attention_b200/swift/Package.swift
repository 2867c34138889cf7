// swift-tools-version: 5.9
// SwiftPM package that keeps the reference's module name and public types (FlashAttention:
// AttentionDescriptor, AttentionKernelDescriptor, AttentionKernel, AttentionKernelType, AttentionOperand,
// GEMMOperandPrecision) over the B200 C ABI (include/mfa_b200.h -> libmfa_b200.so).
// NOTE: no Swift toolchain exists in the build image, so this package is reviewed, not compiled, here;
// all logic lives on the C side of the ABI and the Swift layer is a mechanical forwarding shim.
import PackageDescription

let package = Package(
  name: "FlashAttention",
  products: [.library(name: "FlashAttention", targets: ["FlashAttention"])],
  targets: [
    // C module: Sources/CMFAB200/include/module.modulemap re-exports include/mfa_b200.h
    .systemLibrary(name: "CMFAB200", path: "Sources/CMFAB200"),
    .target(
      name: "FlashAttention",
      dependencies: ["CMFAB200"],
      linkerSettings: [.linkedLibrary("mfa_b200"), .unsafeFlags(["-L../lib", "-Xlinker", "-rpath", "-Xlinker", "../lib"])]),
  ]
)
