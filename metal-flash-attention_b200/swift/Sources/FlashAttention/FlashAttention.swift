//
//  FlashAttention.swift -- drop-in Swift surface over the B200 C ABI.
//
//  Same public names, fields and failure behaviour as the reference package
//  (Sources/FlashAttention/Attention/*.swift), minus everything Metal-specific:
//    * `createSource()` is replaced by `encode(...)`: the kernels are ahead-of-time compiled CUDA, so
//      the compile + bind + dispatch the reference leaves to its caller happens behind the C ABI.
//    * `setFunctionConstants(_:)` fills a plain struct instead of MTLFunctionConstantValues.
//  Where the reference calls fatalError(...), so does this shim (with the C ABI's message).
//
import CMFAB200

/// GEMMOperandPrecision.swift:33-61 (raw values are shared with the C ABI).
public enum GEMMOperandPrecision: UInt16 {
  case FP32 = 0
  case FP16 = 1
  case BF16 = 2
  public var size: Int { self == .FP32 ? 4 : 2 }
  public var name: String { String(cString: mfa_precision_name(mfa_precision_t(UInt32(rawValue)))) }
}

/// AttentionKernelType.swift:8-23
public enum AttentionKernelType: UInt32 {
  case forward = 0
  case backwardQuery = 1
  case backwardKeyValue = 2
}

/// AttentionOperand.swift:8-72; raw values 0...9 are the buffer bindings.
public enum AttentionOperand: UInt32, Hashable, CustomStringConvertible {
  case Q = 0, K, V, O, L, D, dO, dV, dK, dQ, S, P, dP, dS
  public var description: String { String(cString: mfa_operand_name(mfa_operand_t(rawValue))) }
  public var bufferBinding: UInt8? {
    let binding = mfa_operand_buffer_binding(mfa_operand_t(rawValue))
    return binding < 0 ? nil : UInt8(binding)
  }
}

@inline(__always) private func check(_ status: Int32) {
  if status != 0 { fatalError(String(cString: mfa_last_error())) }
}

/// AttentionDescriptor.swift:10-27
public struct AttentionDescriptor {
  public var lowPrecisionInputs: Bool = false
  public var lowPrecisionIntermediates: Bool = false
  public var matrixDimensions: (row: UInt32, column: UInt32, head: UInt16)?
  public var transposeState: (Q: Bool, K: Bool, V: Bool, O: Bool)?
  /// B200 extension: BF16 Q/K/V/dO in memory (nil = the reference's FP16 policy).
  public var inputPrecisionOverride: GEMMOperandPrecision?
  /// B200 extension: number of independent single-head problems stored back to back.
  public var batchCount: UInt32 = 1

  public init() {}

  var c: mfa_attention_descriptor_t {
    var d = mfa_attention_descriptor_t()
    mfa_attention_descriptor_init(&d)
    d.low_precision_inputs = lowPrecisionInputs ? 1 : 0
    d.low_precision_intermediates = lowPrecisionIntermediates ? 1 : 0
    if let m = matrixDimensions {
      d.has_matrix_dimensions = 1
      d.row = m.row; d.column = m.column; d.head = m.head
    }
    if let t = transposeState {
      d.has_transpose_state = 1
      d.transpose_Q = t.Q ? 1 : 0; d.transpose_K = t.K ? 1 : 0
      d.transpose_V = t.V ? 1 : 0; d.transpose_O = t.O ? 1 : 0
    }
    d.input_precision_override = UInt8(inputPrecisionOverride?.rawValue ?? 0)
    d.batch_count = batchCount
    return d
  }

  /// AttentionDescriptor.swift:33-130
  public func kernelDescriptor(type: AttentionKernelType) -> AttentionKernelDescriptor {
    var descriptor = c
    var output = AttentionKernelDescriptor()
    check(mfa_attention_descriptor_kernel_descriptor(&descriptor, mfa_kernel_type_t(type.rawValue), &output.c))
    return output
  }

  /// AttentionDescriptor+Precisions.swift:10-146
  public var memoryPrecisions: [AttentionOperand: GEMMOperandPrecision] {
    var descriptor = c
    var output: [AttentionOperand: GEMMOperandPrecision] = [:]
    for raw in UInt32(0)..<UInt32(MFA_BUFFER_COUNT) {
      var precision = mfa_precision_t(0)
      check(mfa_attention_descriptor_memory_precision(&descriptor, mfa_operand_t(raw), &precision))
      output[AttentionOperand(rawValue: raw)!] = GEMMOperandPrecision(rawValue: UInt16(precision.rawValue))!
    }
    return output
  }

  /// AttentionDescriptor.swift:139-148 (R at index 0, C at index 1).
  public func setFunctionConstants(_ constants: inout mfa_function_constants_t) {
    var descriptor = c
    check(mfa_attention_descriptor_set_function_constants(&descriptor, &constants))
  }
}

/// AttentionKernelDescriptor.swift:7-48 -- a plain, editable value (here: the C struct).
public struct AttentionKernelDescriptor {
  public var c = mfa_attention_kernel_descriptor_t()
  public init() { mfa_attention_kernel_descriptor_init(&c) }
  public var blockDimensions: (parallelization: UInt16, traversal: UInt16, head: UInt16)? {
    c.has_block_dimensions == 0 ? nil : (c.block_parallelization, c.block_traversal, c.block_head)
  }
  public var headDimension: UInt16? { c.has_head_dimension == 0 ? nil : c.head_dimension }
  public var type: AttentionKernelType? { c.type == 0xFF ? nil : AttentionKernelType(rawValue: UInt32(c.type)) }
}

/// AttentionKernel.swift:11-50, 268-363
public final class AttentionKernel {
  let handle: OpaquePointer
  let owned: Bool

  public init(descriptor: AttentionKernelDescriptor) {
    var kd = descriptor.c
    var out: OpaquePointer?
    check(mfa_attention_kernel_create(&kd, &out))
    handle = out!
    owned = true
  }
  /// Library-owned kernel object from the descriptor-keyed cache -- the analogue of
  /// `GEMMKernel.pipelineCache[descriptor]` (GEMMDescriptor+PipelineCache.swift:16-36).
  public init(cached descriptor: AttentionDescriptor, type: AttentionKernelType) {
    var d = descriptor.c
    var out: OpaquePointer?
    check(mfa_attention_kernel_cache_fetch(&d, mfa_kernel_type_t(type.rawValue), &out))
    handle = out!
    owned = false
  }
  deinit { if owned { mfa_attention_kernel_destroy(handle) } }

  public var blockDimensions: (parallelization: UInt16, traversal: UInt16, head: UInt16) {
    var out: (UInt16, UInt16, UInt16) = (0, 0, 0)
    withUnsafeMutablePointer(to: &out) {
      $0.withMemoryRebound(to: UInt16.self, capacity: 3) { check(mfa_attention_kernel_block_dimensions(handle, $0)) }
    }
    return out
  }
  public var threadgroupSize: UInt32 {
    var out: UInt32 = 0
    check(mfa_attention_kernel_threadgroup_size(handle, &out))
    return out
  }
  public var threadgroupMemoryAllocation: UInt32 {
    var out: UInt32 = 0
    check(mfa_attention_kernel_threadgroup_memory_allocation(handle, &out))
    return out
  }

  /// What the reference's callers do by hand around `createSource()` (SquareAttentionTest.swift:240-372):
  /// `buffers[binding]` are DEVICE pointers at AttentionOperand.bufferBinding; `stream` is a cudaStream_t.
  public func encode(constants: mfa_function_constants_t,
                     buffers: [AttentionOperand: UnsafeMutableRawPointer],
                     stream: UnsafeMutableRawPointer? = nil) {
    var table = [UnsafeMutableRawPointer?](repeating: nil, count: Int(MFA_BUFFER_COUNT))
    for (operand, pointer) in buffers {
      guard let binding = operand.bufferBinding else { fatalError("Operand \(operand) has no buffer binding.") }
      table[Int(binding)] = pointer
    }
    var constants = constants
    check(mfa_attention_kernel_encode(handle, &constants, &table, stream))
  }
}
