// fmt/core.h — SHIM. TEST INFRASTRUCTURE ONLY (oracle/ref). The reference fetches {fmt} 7.1.3 from GitHub
// (external/CMakeLists.txt:23-27); it is not vendored and there is no network. Only log / exit messages go through it —
// no arithmetic of the render path — so this is a minimal stand-in: "{}" and "{name:spec}" are replaced by the streamed
// argument, format specs are ignored.
#pragma once
#include <algorithm>
#include <cstdio>
#include <functional>
#include <iostream>
#include <map>
#include <memory>
#include <unordered_map>
#include <ctime>
#include <sstream>
#include <string>
#include <string_view>
#include <utility>
#include <vector>

namespace fmt
{
    template<typename T>
    struct named_arg
    {
        const char *name;
        const T    &value;
    };
    template<typename T>
    named_arg<T> arg(const char *name, const T &value)
    {
        return named_arg<T> { name, value };
    }
    namespace detail
    {
        template<typename T>
        void put(std::vector<std::pair<std::string, std::string>> &out, const T &v)
        {
            std::ostringstream s;
            s << v;
            out.push_back({ "", s.str() });
        }
        template<typename T>
        void put(std::vector<std::pair<std::string, std::string>> &out, const named_arg<T> &v)
        {
            std::ostringstream s;
            s << v.value;
            out.push_back({ v.name, s.str() });
        }
    }    // namespace detail
    template<typename... A>
    std::string format(std::string_view f, const A &...a)
    {
        std::vector<std::pair<std::string, std::string>> args;
        (detail::put(args, a), ...);
        std::string out;
        size_t      next = 0;
        for (size_t i = 0; i < f.size(); i++)
        {
            if (f[i] == '{')
            {
                const size_t e = f.find('}', i);
                if (e == std::string_view::npos) break;
                std::string key(f.substr(i + 1, e - i - 1));
                key = key.substr(0, key.find(':'));
                bool done = false;
                if (!key.empty())
                    for (auto &kv : args)
                        if (kv.first == key) out += kv.second, done = true;
                if (!done && next < args.size()) out += args[next++].second;
                i = e;
            }
            else
                out += f[i];
        }
        return out;
    }
    template<typename... A>
    void print(std::string_view f, const A &...a)
    {
        std::fputs(format(f, a...).c_str(), stdout);
    }
    inline std::tm localtime(std::time_t t)
    {
        std::tm r {};
        localtime_r(&t, &r);
        return r;
    }
    enum class color { red, yellow, white, blue_violet };
}    // namespace fmt
