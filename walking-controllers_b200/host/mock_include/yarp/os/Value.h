// MOCK: yarp::os::Value / Bottle / Searchable, genuine members only (YARP 3.1)
#pragma once
#include <map>
#include <memory>
#include <string>
#include <vector>
namespace yarp { namespace os {
class Bottle;
class Value {
    int kind_ = 0; // 0 null, 1 int, 2 double, 3 string, 4 list
    int i_ = 0; double d_ = 0.0; std::string s_; std::shared_ptr<Bottle> l_;
public:
    Value() {}
    Value(int x, bool isVocab = false) : kind_(1), i_(x), d_(x) { (void)isVocab; }
    Value(double x) : kind_(2), i_((int)x), d_(x) {}
    Value(const std::string& str, bool isVocab = false) : kind_(3), s_(str) { (void)isVocab; }
    bool isNull() const { return kind_ == 0; }
    bool isInt() const { return kind_ == 1; }
    bool isDouble() const { return kind_ == 2; }
    bool isString() const { return kind_ == 3; }
    bool isList() const { return kind_ == 4; }
    bool isBool() const { return kind_ == 1; }
    int asInt() const { return kind_ == 2 ? (int)d_ : i_; }
    double asDouble() const { return kind_ == 1 ? (double)i_ : d_; }
    bool asBool() const { return asInt() != 0; }
    std::string asString() const { return s_; }
    Bottle* asList() const { return l_.get(); }
    static Value* makeList() { Value* v = new Value(); v->kind_ = 4; v->l_ = std::make_shared<Bottle>(); return v; }
    static Value& getNullValue() { static Value n; return n; }
};
class Searchable {
public:
    virtual ~Searchable() {}
    virtual bool check(const std::string& key) const = 0;
    virtual Value& find(const std::string& key) const = 0;
    virtual bool isNull() const { return false; }
    virtual Value check(const std::string& key, const Value& fallback, const std::string& comment = "") const
    { (void)comment; return check(key) ? find(key) : fallback; }
};
class Bottle : public Searchable {
    std::vector<std::shared_ptr<Value>> items_;
public:
    int size() const { return (int)items_.size(); }
    Value& get(int index) const { return (index < 0 || index >= size()) ? Value::getNullValue() : *items_[index]; }
    void addInt(int x) { items_.push_back(std::make_shared<Value>(x)); }
    void addDouble(double x) { items_.push_back(std::make_shared<Value>(x)); }
    void addString(const std::string& str) { items_.push_back(std::make_shared<Value>(str)); }
    Bottle& addList() { std::shared_ptr<Value> v(Value::makeList()); items_.push_back(v); return *v->asList(); }
    bool check(const std::string& key) const override { return !find(key).isNull(); }
    Value& find(const std::string& key) const override
    { // "(key value...)" sub-lists
        for (const auto& it : items_) if (it->isList() && it->asList()->size() >= 2 && it->asList()->get(0).asString() == key) return it->asList()->get(1);
        return Value::getNullValue();
    }
    bool isNull() const override { return items_.empty(); }
};
class Property : public Searchable {
    mutable std::map<std::string, std::shared_ptr<Value>> m_;
public:
    void put(const std::string& key, const Value& value) { m_[key] = std::make_shared<Value>(value); }
    void put(const std::string& key, Value* value) { m_[key] = std::shared_ptr<Value>(value); }
    bool check(const std::string& key) const override { return m_.count(key) != 0; }
    Value& find(const std::string& key) const override { auto it = m_.find(key); return it == m_.end() ? Value::getNullValue() : *it->second; }
    bool isNull() const override { return m_.empty(); }
};
}} // namespace yarp::os
