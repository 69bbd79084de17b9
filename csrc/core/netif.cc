#include "core/netif.h"

#include <ifaddrs.h>
#include <net/if.h>
#include <string.h>

#include <set>

namespace bnet {

IfFilter IfFilter::parse(const std::string& spec) {
  IfFilter f;
  std::string s = spec;
  if (!s.empty() && s[0] == '^') {
    f.exclude = true;
    s = s.substr(1);
  }
  if (!s.empty() && s[0] == '=') {
    f.exact = true;
    s = s.substr(1);
  }
  size_t pos = 0;
  while (pos <= s.size()) {
    size_t c = s.find(',', pos);
    if (c == std::string::npos) c = s.size();
    if (c > pos) f.names.push_back(s.substr(pos, c - pos));
    pos = c + 1;
  }
  return f;
}

bool IfFilter::accepts(const std::string& ifname) const {
  bool hit = false;
  for (const auto& n : names) {
    if (exact ? (ifname == n) : (ifname.compare(0, n.size(), n) == 0)) {
      hit = true;
      break;
    }
  }
  if (names.empty()) return true;
  return exclude ? !hit : hit;
}

static std::vector<NetIf> scan(const IfFilter& filt, int family, bool allow_loopback) {
  std::vector<NetIf> out;
  ifaddrs* ifa = nullptr;
  if (getifaddrs(&ifa) != 0) return out;
  std::set<std::string> seen;
  for (ifaddrs* it = ifa; it; it = it->ifa_next) {
    if (!it->ifa_addr || !it->ifa_name) continue;
    int af = it->ifa_addr->sa_family;
    if (af != AF_INET && af != AF_INET6) continue;
    if (family >= 0 && af != family) continue;
    if (!(it->ifa_flags & IFF_UP)) continue;
    bool lo = (it->ifa_flags & IFF_LOOPBACK) != 0;
    if (lo && !allow_loopback) continue;
    std::string name = it->ifa_name;
    if (name.size() >= IFNAMSIZ) continue;  // the reference asserts here (utils.rs:64)
    if (!filt.accepts(name)) continue;
    if (af == AF_INET6) {
      // link-local v6 needs a scope id to be connectable; skip unless it has one
      const sockaddr_in6* s6 = (const sockaddr_in6*)it->ifa_addr;
      if (IN6_IS_ADDR_LINKLOCAL(&s6->sin6_addr) && s6->sin6_scope_id == 0) continue;
    }
    if (!seen.insert(name).second) continue;  // first address per interface wins
    NetIf n;
    n.name = name;
    memset(&n.addr, 0, sizeof(n.addr));
    memcpy(&n.addr, it->ifa_addr, af == AF_INET ? sizeof(sockaddr_in) : sizeof(sockaddr_in6));
    n.pci_path = net_if_pci_path(name);
    n.speed_mbps = net_if_speed_mbps(name);
    n.loopback = lo;
    out.push_back(n);
  }
  freeifaddrs(ifa);
  return out;
}

std::vector<NetIf> find_interfaces(const char* ifname_spec, int family) {
  if (family == -2) family = (int)env_plain_int("NCCL_SOCKET_FAMILY", -1);
  const char* env = ifname_spec ? ifname_spec : env_plain("NCCL_SOCKET_IFNAME");
  bool user_spec = env != nullptr;
  IfFilter filt = IfFilter::parse(user_spec ? env : "^docker,lo");
  // a user who names `lo` explicitly gets it; otherwise loopback is skipped first
  bool user_wants_lo = false;
  if (user_spec && !filt.exclude)
    for (auto& n : filt.names)
      if (n.compare(0, 2, "lo") == 0) user_wants_lo = true;
  std::vector<NetIf> devs = scan(filt, family, user_wants_lo);
  if (devs.empty() && !user_spec) {
    // isolated box: only loopback exists -> serve it
    devs = scan(IfFilter::parse("^docker"), family, true);
    if (!devs.empty()) BNET_DEBUG("no external NIC found, falling back to %s", devs[0].name.c_str());
  }
  return devs;
}

}  // namespace bnet
