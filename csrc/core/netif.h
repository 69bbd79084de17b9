// NIC discovery with NCCL's NCCL_SOCKET_IFNAME / NCCL_SOCKET_FAMILY syntax.
// Parity: reference src/utils.rs:32-130 (find_interfaces).  Differences by
// design: '=' means exact match (NCCL semantics; the reference prefix-matches),
// and when nothing but loopback exists we fall back to `lo` the way NCCL's own
// socket transport does, so an isolated box still gets one device.
#pragma once
#include <string>
#include <vector>

#include "core/common.h"

namespace bnet {

struct NetIf {
  std::string name;
  SockAddr addr;
  std::string pci_path;
  int speed_mbps;
  bool loopback;
};

struct IfFilter {
  bool exclude = false;   // '^' prefix
  bool exact = false;     // '=' prefix
  std::vector<std::string> names;
  static IfFilter parse(const std::string& spec);
  bool accepts(const std::string& ifname) const;
};

// All usable interfaces (one entry per interface name, first address wins).
// `ifname_spec`/`family` default to the NCCL_* environment.
std::vector<NetIf> find_interfaces(const char* ifname_spec = nullptr, int family = -2);

}  // namespace bnet
