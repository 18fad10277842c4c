// oracle/refbuild/boost/program_options.hpp - TEST INFRASTRUCTURE, not product code.
//
// Boost is not installed in this image.  The reference needs boost::program_options in exactly one
// function, SyphaEnvironment::readInputArguments (src/sypha_environment.cpp:104-246), i.e. for its CLI
// (src/main.cpp).  This header implements the small subset that function uses - options_description with
// chained add_options(), value<T>(&target)->default_value(v), bool_switch, parse_command_line for
// "--name value" / "--name=value" / bare switches, variables_map::count / operator[] / as<T>() - so that the
// UNMODIFIED reference CLI builds and runs as the side-by-side baseline (oracle/Makefile -> oracle/_ref/).
#ifndef SB200_REFBUILD_BOOST_PROGRAM_OPTIONS_HPP
#define SB200_REFBUILD_BOOST_PROGRAM_OPTIONS_HPP
#include <map>
#include <memory>
#include <ostream>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

namespace boost
{
namespace program_options
{

struct error : std::runtime_error
{
    explicit error(const std::string &w) : std::runtime_error(w) {}
};

class value_semantic
{
  public:
    virtual ~value_semantic() = default;
    virtual bool is_switch() const = 0;
    virtual bool has_default() const = 0;
    virtual void apply_default() = 0;
    virtual void parse(const std::string &text) = 0;
    virtual const void *stored() const = 0;
};

namespace detail
{
template <class T> inline T from_text(const std::string &s)
{
    std::istringstream is(s);
    T v{};
    is >> v;
    if (is.fail()) throw error("the argument ('" + s + "') is invalid");
    return v;
}
template <> inline std::string from_text<std::string>(const std::string &s) { return s; }
template <> inline bool from_text<bool>(const std::string &s)
{
    if (s == "1" || s == "true" || s == "on" || s == "yes") return true;
    if (s == "0" || s == "false" || s == "off" || s == "no") return false;
    throw error("the argument ('" + s + "') is invalid");
}
} // namespace detail

template <class T> class typed_value : public value_semantic
{
  public:
    explicit typed_value(T *target, bool sw = false) : target_(target), switch_(sw) {}
    typed_value *default_value(const T &v)
    {
        default_ = v;
        has_default_ = true;
        return this;
    }
    bool is_switch() const override { return switch_; }
    bool has_default() const override { return has_default_; }
    void apply_default() override { set(default_); }
    void parse(const std::string &text) override { set(switch_ ? detail::from_text<T>("true") : detail::from_text<T>(text)); }
    const void *stored() const override { return &value_; }

  private:
    void set(const T &v)
    {
        value_ = v;
        if (target_) *target_ = v;
    }
    T *target_;
    bool switch_;
    bool has_default_ = false;
    T default_{};
    T value_{};
};

template <class T> typed_value<T> *value(T *target = nullptr) { return new typed_value<T>(target); }
inline typed_value<bool> *bool_switch(bool *target = nullptr)
{
    typed_value<bool> *v = new typed_value<bool>(target, true);
    v->default_value(false);
    return v;
}

struct option_description
{
    std::string name, help;
    std::shared_ptr<value_semantic> semantic; // null: presence-only option ("help")
};

class options_description;
class options_description_easy_init
{
  public:
    explicit options_description_easy_init(options_description *o) : owner_(o) {}
    options_description_easy_init &operator()(const char *name, const char *help);
    options_description_easy_init &operator()(const char *name, value_semantic *s, const char *help);

  private:
    options_description *owner_;
};

class options_description
{
  public:
    explicit options_description(const std::string &caption = "") : caption_(caption) {}
    options_description_easy_init add_options() { return options_description_easy_init(this); }
    const option_description *find(const std::string &name) const
    {
        for (const auto &o : options_)
            if (o.name == name) return &o;
        return nullptr;
    }
    std::string caption_;
    std::vector<option_description> options_;
};

inline options_description_easy_init &options_description_easy_init::operator()(const char *name, const char *help)
{
    owner_->options_.push_back({name, help, nullptr});
    return *this;
}
inline options_description_easy_init &options_description_easy_init::operator()(const char *name, value_semantic *s,
                                                                                 const char *help)
{
    owner_->options_.push_back({name, help, std::shared_ptr<value_semantic>(s)});
    return *this;
}

inline std::ostream &operator<<(std::ostream &os, const options_description &d)
{
    os << d.caption_ << ":\n";
    for (const auto &o : d.options_) os << "  --" << o.name << "\t" << o.help << "\n";
    return os;
}

class variable_value
{
  public:
    variable_value() = default;
    explicit variable_value(std::shared_ptr<value_semantic> s) : semantic_(std::move(s)) {}
    template <class T> const T &as() const
    {
        if (!semantic_) throw error("option has no value");
        return *static_cast<const T *>(semantic_->stored());
    }

  private:
    std::shared_ptr<value_semantic> semantic_;
};

struct parsed_options
{
    const options_description *description;
    std::vector<std::pair<std::string, std::string>> given; // (name, text); text empty for switches
};

inline parsed_options parse_command_line(int argc, const char *const argv[], const options_description &desc)
{
    parsed_options out{&desc, {}};
    for (int i = 1; i < argc; ++i)
    {
        std::string tok = argv[i];
        if (tok.rfind("--", 0) != 0) throw error("too many positional options have been specified on the command line");
        tok = tok.substr(2);
        std::string text;
        bool has_text = false;
        const size_t eq = tok.find('=');
        if (eq != std::string::npos)
        {
            text = tok.substr(eq + 1);
            tok = tok.substr(0, eq);
            has_text = true;
        }
        const option_description *o = desc.find(tok);
        if (!o) throw error("unrecognised option '--" + tok + "'");
        const bool takes_value = o->semantic && !o->semantic->is_switch();
        if (takes_value && !has_text)
        {
            if (i + 1 >= argc) throw error("the required argument for option '--" + tok + "' is missing");
            text = argv[++i];
        }
        out.given.emplace_back(tok, text);
    }
    return out;
}

class variables_map
{
  public:
    size_t count(const std::string &name) const { return values_.count(name); }
    const variable_value &operator[](const std::string &name) const
    {
        static const variable_value empty;
        auto it = values_.find(name);
        return it == values_.end() ? empty : it->second;
    }
    std::map<std::string, variable_value> values_;
};

inline void store(const parsed_options &po, variables_map &vm)
{
    for (const auto &g : po.given)
    {
        const option_description *o = po.description->find(g.first);
        if (o->semantic) o->semantic->parse(g.second);
        vm.values_[g.first] = variable_value(o->semantic);
    }
    for (const auto &o : po.description->options_)
        if (o.semantic && o.semantic->has_default() && !vm.values_.count(o.name))
        {
            o.semantic->apply_default();
            vm.values_[o.name] = variable_value(o.semantic);
        }
}

inline void notify(variables_map &) {}

} // namespace program_options
} // namespace boost
#endif
