// Ordered basic / non-basic column lists (reference src/Basis.h:18-34, src/Basis.cpp).
// ORDER IS SIGNIFICANT: the device keys entering candidates by POSITION in the non-basic list and
// leaving candidates by ROW in the basic list, and a pivot swaps basic[row] <-> non_basic[pos]
// in place (src/Basis.cpp:48-53). The lists travel through the C ABI as int64 arrays.
#pragma once

#include "model.h"

#include <algorithm>
#include <string>
#include <vector>

namespace simplex
{
    struct Basis
    {
        [[nodiscard]] const std::vector<Index> & get_basic_columns() const { return m_basic; }
        [[nodiscard]] const std::vector<Index> & get_non_basic_columns() const { return m_non_basic; }
        // raw access for the marshal into sb200_run (lists are updated in place by the device run)
        std::vector<Index> & basic_columns() { return m_basic; }
        std::vector<Index> & non_basic_columns() { return m_non_basic; }

        void insert(Index var_id)
        {
            m_basic.push_back(var_id);
            drop(m_non_basic, var_id);
        }
        void insert_to_non_basic(Index var_id) { m_non_basic.push_back(var_id); }
        void remove(Index var_id)
        {
            if (!drop(m_basic, var_id)) drop(m_non_basic, var_id);
        }
        void clear()
        {
            m_basic.clear();
            m_non_basic.clear();
        }
        bool contains(Index var_id) const { return std::find(m_basic.begin(), m_basic.end(), var_id) != m_basic.end(); }
        void update(Index entering_index, Index leaving_index)
        {
            std::swap(m_basic[static_cast<size_t>(leaving_index)], m_non_basic[static_cast<size_t>(entering_index)]);
        }
        std::string show() const
        {
            std::string text;
            for (Index col : m_basic) text += std::to_string(col) + " ";
            return text;
        }

    private:
        static bool drop(std::vector<Index> & list, Index var_id)
        {
            const auto it = std::find(list.begin(), list.end(), var_id);
            if (it == list.end()) return false;
            list.erase(it);
            return true;
        }
        std::vector<Index> m_basic, m_non_basic;
    };
}  // namespace simplex
