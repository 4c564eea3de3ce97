// psi/seed.hpp -- the seed hit record handed to callbacks.
//
// Mirrors psi::Seed<> of the reference (include/psi/seed.hpp:32-46): six
// size_t fields; the CLI writes the first four (src/psikt.cpp:172-181).
#ifndef PSI_B200_PSI_SEED_HPP
#define PSI_B200_PSI_SEED_HPP

#include <cstddef>

namespace psi {

template <typename TId = std::size_t, typename TOffset = std::size_t>
class Seed {
 public:
  typedef TId id_type;
  typedef TOffset offset_type;
  id_type node_id;          // graph node id (the id space the finder was given, internal ids for psikt)
  offset_type node_offset;  // offset of the seed's first base in the node label
  id_type read_id;          // 0-based ordinal of the read over all chunks
  offset_type read_offset;  // offset of the seed in the read
  offset_type match_len;    // = seed length
  offset_type gocc;         // occurrence count (0 when not tracked)
};

}  // namespace psi
#endif
