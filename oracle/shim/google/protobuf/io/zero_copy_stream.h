// TEST INFRASTRUCTURE ONLY (oracle/): protobuf is absent from this image and
// file I/O is out of the hot path; the reference headers only need the names.
#ifndef VBX_ORACLE_SHIM_PROTOBUF_STUB_
#define VBX_ORACLE_SHIM_PROTOBUF_STUB_
namespace google {
namespace protobuf {
class MessageLite {};
class Message : public MessageLite {};
}  // namespace protobuf
}  // namespace google
#endif
