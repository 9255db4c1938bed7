// Stand-in for <tinyxml.h>: a small in-memory document tree with the members the reference's XMLUtils.h and
// RadioReceiver.cpp (LoadChannelData / SaveChannelData) call.  "Files" are kept in a process-wide map (SaveFile stores
// a copy of the tree under its path, LoadFile fetches it), so the add-on's settings round-trip without a file system
// or an XML parser.  TEST INFRASTRUCTURE ONLY; written from the names the reference uses, not from TinyXML.
#pragma once
#include <map>
#include <memory>
#include <string>
#include <vector>

class TiXmlElement;
class TiXmlDeclaration;
class TiXmlDocument;

class TiXmlNode
{
public:
  enum NodeType { TINYXML_DOCUMENT, TINYXML_ELEMENT, TINYXML_COMMENT, TINYXML_UNKNOWN, TINYXML_TEXT, TINYXML_DECLARATION };
  explicit TiXmlNode(NodeType t = TINYXML_UNKNOWN, const std::string& v = "") : m_type(t), m_value(v) {}
  TiXmlNode(const TiXmlNode& o) : m_type(o.m_type), m_value(o.m_value)
  {
    for (const auto& c : o.m_children)
      m_children.emplace_back(c->Clone());
  }
  TiXmlNode& operator=(const TiXmlNode& o)
  {
    if (this != &o)
    {
      m_type = o.m_type;
      m_value = o.m_value;
      m_children.clear();
      for (const auto& c : o.m_children)
        m_children.emplace_back(c->Clone());
    }
    return *this;
  }
  virtual ~TiXmlNode() = default;
  virtual TiXmlNode* Clone() const { return new TiXmlNode(*this); }

  int Type() const { return m_type; }
  const char* Value() const { return m_value.c_str(); }
  const std::string& ValueStr() const { return m_value; }
  void SetValue(const std::string& v) { m_value = v; }

  TiXmlNode* InsertEndChild(const TiXmlNode& n)
  {
    m_children.emplace_back(n.Clone());
    return m_children.back().get();
  }
  const TiXmlNode* FirstChild() const { return m_children.empty() ? nullptr : m_children.front().get(); }
  TiXmlNode* FirstChild() { return m_children.empty() ? nullptr : m_children.front().get(); }
  const TiXmlNode* FirstChild(const std::string& tag) const
  {
    for (const auto& c : m_children)
      if (c->m_value == tag)
        return c.get();
    return nullptr;
  }
  TiXmlNode* FirstChild(const std::string& tag) { return const_cast<TiXmlNode*>(static_cast<const TiXmlNode*>(this)->FirstChild(tag)); }
  inline const TiXmlElement* FirstChildElement(const std::string& tag) const;
  inline TiXmlElement* FirstChildElement(const std::string& tag);
  const TiXmlNode* IterateChildren(const TiXmlNode* prev) const
  {
    if (!prev)
      return FirstChild();
    for (size_t i = 0; i + 1 < m_children.size(); ++i)
      if (m_children[i].get() == prev)
        return m_children[i + 1].get();
    return nullptr;
  }
  TiXmlNode* IterateChildren(const TiXmlNode* prev) { return const_cast<TiXmlNode*>(static_cast<const TiXmlNode*>(this)->IterateChildren(prev)); }
  inline const TiXmlDeclaration* ToDeclaration() const;
  inline const TiXmlElement* ToElement() const;

protected:
  NodeType m_type;
  std::string m_value;
  std::vector<std::unique_ptr<TiXmlNode>> m_children;
};

class TiXmlElement : public TiXmlNode
{
public:
  explicit TiXmlElement(const std::string& name) : TiXmlNode(TINYXML_ELEMENT, name) {}
  TiXmlNode* Clone() const override { return new TiXmlElement(*this); }
  void SetAttribute(const std::string& name, int value) { m_attr[name] = std::to_string(value); }
  void SetAttribute(const std::string& name, const std::string& value) { m_attr[name] = value; }
  const char* Attribute(const std::string& name) const
  {
    auto it = m_attr.find(name);
    return it == m_attr.end() ? nullptr : it->second.c_str();
  }

private:
  std::map<std::string, std::string> m_attr;
};

class TiXmlText : public TiXmlNode
{
public:
  explicit TiXmlText(const std::string& text) : TiXmlNode(TINYXML_TEXT, text) {}
  TiXmlNode* Clone() const override { return new TiXmlText(*this); }
};

class TiXmlDeclaration : public TiXmlNode
{
public:
  TiXmlDeclaration() : TiXmlNode(TINYXML_DECLARATION) {}
  const char* Encoding() const { return "UTF-8"; }
  TiXmlNode* Clone() const override { return new TiXmlDeclaration(*this); }
};

class TiXmlDocument : public TiXmlNode
{
public:
  TiXmlDocument() : TiXmlNode(TINYXML_DOCUMENT) {}
  TiXmlNode* Clone() const override { return new TiXmlDocument(*this); }
  static std::map<std::string, TiXmlDocument>& Store()
  {
    static std::map<std::string, TiXmlDocument> files;
    return files;
  }
  bool LoadFile(const std::string& path)
  {
    auto it = Store().find(path);
    if (it == Store().end())
      return false;
    *static_cast<TiXmlNode*>(this) = it->second;
    return true;
  }
  bool SaveFile(const std::string& path) const
  {
    Store()[path] = *this;
    return true;
  }
  const TiXmlElement* RootElement() const { return ToFirstElement(); }
  TiXmlElement* RootElement() { return const_cast<TiXmlElement*>(ToFirstElement()); }

private:
  const TiXmlElement* ToFirstElement() const
  {
    for (const auto& c : m_children)
      if (c->Type() == TINYXML_ELEMENT)
        return static_cast<const TiXmlElement*>(c.get());
    return nullptr;
  }
};

inline const TiXmlElement* TiXmlNode::FirstChildElement(const std::string& tag) const
{
  for (const auto& c : m_children)
    if (c->m_type == TINYXML_ELEMENT && c->m_value == tag)
      return static_cast<const TiXmlElement*>(c.get());
  return nullptr;
}
inline TiXmlElement* TiXmlNode::FirstChildElement(const std::string& tag)
{
  return const_cast<TiXmlElement*>(static_cast<const TiXmlNode*>(this)->FirstChildElement(tag));
}
inline const TiXmlDeclaration* TiXmlNode::ToDeclaration() const
{
  return m_type == TINYXML_DECLARATION ? static_cast<const TiXmlDeclaration*>(this) : nullptr;
}
inline const TiXmlElement* TiXmlNode::ToElement() const
{
  return m_type == TINYXML_ELEMENT ? static_cast<const TiXmlElement*>(this) : nullptr;
}
